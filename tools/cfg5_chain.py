"""BASELINE.json configs[4]: `colorlut(33^3) ! hsvfilter(cfg2)` on 7680x4320 RGBA, one 64-frame
batch sharded frame-parallel (round-robin, no collective) over the GPUs of the box — STRONG scaling:
the 64 frames are the whole job at every N.  Prints one JSON line (rank 0).

  python tools/cfg5_chain.py                                   # 1 GPU
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/cfg5_chain.py
  python tools/cfg5_chain.py --single-process --gpus N         # one process, b200vf_group over N GPUs
  python tools/cfg5_chain.py --cpu                             # oracle port on the host cores
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

W, H, N_FRAMES, LUT_N = 7680, 4320, 64, 33
CFG2 = (37.5, 1.2, 0.05, 0.9, 0.02)


def gpu_main(args):
    import torch
    import gst_plugins_rs_b200 as g
    from gst_plugins_rs_b200 import frames, sharding
    from gst_plugins_rs_b200.api import frame_array, frame_of
    rank, local, world = sharding.world()
    torch.cuda.set_device(local)
    use_dist = sharding.init_process_group("nccl", torch.device("cuda", local))
    ctx = g.Context(local)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx.set_lut_from_cube(g.parse_cube(frames.cube_text_3d(LUT_N)))
    mine = sharding.frames_for_rank(N_FRAMES, rank, world)
    uniq = [torch.from_numpy(frames.frame_of_class(args.content, W, H, i).reshape(-1)).cuda()
            for i in range(4)]
    src = [uniq[i % 4].clone() for i in mine]          # this rank's share of the 64 frames
    dst = [torch.empty_like(t) for t in src]
    fin = frame_array([frame_of(t, W, H, "RGBA") for t in src])
    fout = frame_array([frame_of(t, W, H, "RGBA") for t in dst])
    p = g.HsvFilterParams(*CFG2)

    def fused():
        ctx.chain_lut_hsv_batch(fin, fout, p)

    def two_pass():                                     # the two elements back to back
        ctx.colorlut_batch(fin, fout)
        ctx.hsvfilter_batch(fout, p)

    out = {}
    # "fused": the library's default policy — compute kernels first, then the chain's function
    # table once it measures faster; "fused_compute_kernel": the fused per-pixel kernel only
    for name, fn, hsv_path in (("fused", fused, 0), ("fused_compute_kernel", fused, 1),
                               ("two_elements", two_pass, 0)):
        ctx.set_option("hsv.path", hsv_path)
        for _ in range(8):  # lets the auto policy build and time both ways before the clock starts
            fn()
        if use_dist:
            sharding.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        if use_dist:
            sharding.barrier()
        torch.cuda.synchronize()
        ms = sharding.max_over_ranks(e0.elapsed_time(e1), "cuda") / args.steps
        out[name] = {"ms_per_64_frame_batch": ms, "frames_per_s": N_FRAMES / (ms / 1e3)}
    if rank == 0:
        print(json.dumps({"config": "colorlut(33^3) ! hsvfilter 7680x4320 RGBA, 64-frame batch, "
                                    "frame-parallel", "content": args.content, "n_gpus": world,
                          "scaling": "strong", **out}))
    if use_dist:
        import torch.distributed as dist
        dist.destroy_process_group()


def group_main(args):
    """The same job from ONE process: b200vf_group, frame i on (and processed by) device i mod N."""
    import torch
    import gst_plugins_rs_b200 as g
    from gst_plugins_rs_b200 import frames
    from gst_plugins_rs_b200.api import frame_array, frame_of
    devs = list(range(args.gpus))
    grp = g.Group(devs)
    grp.set_lut_from_cube(g.parse_cube(frames.cube_text_3d(LUT_N)))
    streams = [torch.cuda.ExternalStream(grp.member(m).get_stream(), device=f"cuda:{d}") for m, d in enumerate(devs)]
    host = [torch.from_numpy(frames.frame_of_class(args.content, W, H, i).reshape(-1)) for i in range(4)]
    seeds = {(d, k): host[k].to(f"cuda:{d}") for d in devs for k in range(4)}
    src = [seeds[(devs[i % len(devs)], i % 4)].clone() for i in range(N_FRAMES)]
    dst = [torch.empty_like(t) for t in src]
    fin = frame_array([frame_of(t, W, H, "RGBA") for t in src])
    fout = frame_array([frame_of(t, W, H, "RGBA") for t in dst])
    p = g.HsvFilterParams(*CFG2)
    out = {}
    for name, hsv_path in (("fused", 0), ("fused_compute_kernel", 1)):
        grp.set_option("hsv.path", hsv_path)
        for _ in range(8):
            grp.chain_lut_hsv_batch(fin, fout, p)
            grp.synchronize()
        ev = [(torch.cuda.Event(True), torch.cuda.Event(True)) for _ in streams]
        for (e0, _), st in zip(ev, streams):
            e0.record(st)
        for _ in range(args.steps):
            grp.chain_lut_hsv_batch(fin, fout, p)
        for (_, e1), st in zip(ev, streams):
            e1.record(st)
        grp.synchronize()
        ms = max(e0.elapsed_time(e1) for e0, e1 in ev) / args.steps
        out[name] = {"ms_per_64_frame_batch": ms, "frames_per_s": N_FRAMES / (ms / 1e3)}
    print(json.dumps({"config": "colorlut(33^3) ! hsvfilter 7680x4320 RGBA, 64-frame batch, frame-parallel",
                      "content": args.content, "n_gpus": args.gpus, "scaling": "strong",
                      "process_model": "one process, b200vf_group", **out}))
    grp.close()


def cpu_main(args):
    import numpy as np
    import oracle
    from gst_plugins_rs_b200 import frames
    oracle.build()
    cores = os.cpu_count() or 1
    n = min(N_FRAMES, cores)                            # bounded sample: one frame per core
    lut = oracle.Lut(text=frames.cube_text_3d(LUT_N))
    uniq = [frames.frame_of_class(args.content, W, H, i).reshape(-1) for i in range(2)]
    src = [uniq[i % 2].copy() for i in range(n)]
    dst = [np.empty_like(s) for s in src]
    t0 = time.perf_counter()
    oracle.colorlut_frames_mt(lut, src, dst, W, H, "RGBA", cores)
    oracle.hsvfilter_frames_mt(dst, W, H, "RGBA", CFG2, cores)
    dt = time.perf_counter() - t0
    print(json.dumps({"config": "colorlut(33^3) ! hsvfilter 7680x4320 RGBA (CPU port of the reference)",
                      "cores": cores, "frames_timed": n, "frames_per_s": n / dt}))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--content", default="grad")
    ap.add_argument("--single-process", action="store_true")
    ap.add_argument("--gpus", type=int, default=1)
    a = ap.parse_args()
    cpu_main(a) if a.cpu else group_main(a) if a.single_process else gpu_main(a)
