"""Summarise an .ncu-rep (read here, no GPU): key metrics + per-opcode instruction mix.
usage: python tools/ncu_summary.py <rep> <pixels-per-launch> [out.md]"""
import collections
import csv
import re
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum",
        "lts__t_sectors_op_read.sum", "lts__t_sectors_srcunit_tex.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_l1tex2xbar_write_bytes.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum",
        "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts_mem_lg.sum",
        "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active",
        "l1tex__lsu_writeback_active_mem_lg.sum",
        "l1tex__f_wavefronts.sum", "l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_not_selected_per_warp_active.pct",
        "smsp__warp_issue_stalled_wait_per_warp_active.pct",
        "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_tex_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct",
        ]


def main():
    rep, px = sys.argv[1], float(sys.argv[2])
    out = open(sys.argv[3], "w") if len(sys.argv) > 3 else sys.stdout
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    d = data[0]
    print(f"# {rep}\n", file=out)
    print("kernel:", d[hdr.index("Kernel Name")], "\n", file=out)
    print("| metric | value | unit |\n|---|---|---|", file=out)
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"| {w} | {d[i]} | {units[i]} |", file=out)
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    h = rows[1]
    i_src, i_ex, i_s = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
    ops, samp, total, n = collections.Counter(), collections.Counter(), 0, 0
    for r in rows[2:]:
        if len(r) < 10 or r[0] in ("Kernel Name", "Address"):
            if n:
                break
            continue
        n += 1
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[i_src].strip())
        op = m.group(2) if m else r[i_src]
        ops[op] += int(r[i_ex])
        samp[op] += int(r[i_s])
        total += int(r[i_ex])
    print(f"\nwarp-instructions executed: {total}  = {total * 32 / px:.2f} per pixel\n", file=out)
    print("| opcode | per pixel | stall samples |\n|---|---|---|", file=out)
    for op, c in ops.most_common(28):
        print(f"| {op} | {c * 32 / px:.2f} | {samp[op]} |", file=out)


if __name__ == "__main__":
    main()
