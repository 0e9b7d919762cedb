"""Where does a host-frame call spend its time?  (dev aid)  Reports, per chunk size: frames/s, host time
blocked on slot events, host time enqueueing, and raw chunked cudaMemcpyAsync both ways for comparison."""
import os
import sys
os.environ["B200VF_ALLOW_DEBUG_MODES"] = "1"  # modes 1-3 skip the kernels (wrong pixels): analysis only
import time
import torch
sys.path.insert(0, ".")
import gst_plugins_rs_b200 as g
from gst_plugins_rs_b200 import frames
from gst_plugins_rs_b200.api import frame_array, frame_of

w, h = 3840, 2160
fb = w * h * 4


def raw_chunked(chunk, nframes=8, iters=6):
    hin = torch.empty(fb * nframes, dtype=torch.uint8).pin_memory()
    hout = torch.empty(fb * nframes, dtype=torch.uint8).pin_memory()
    din = torch.empty(fb * nframes, dtype=torch.uint8, device="cuda")
    dout = torch.empty(fb * nframes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run():
        for off in range(0, fb * nframes, chunk):
            n = min(chunk, fb * nframes - off)
            with torch.cuda.stream(s1):
                din[off:off + n].copy_(hin[off:off + n], non_blocking=True)
            with torch.cuda.stream(s2):
                hout[off:off + n].copy_(dout[off:off + n], non_blocking=True)
    run()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        run()
    torch.cuda.synchronize()
    return fb * nframes * iters / (time.perf_counter() - t0) / 1e9


def main():
    for chunk in (1 << 20, 4 << 20, 16 << 20):
        print("raw chunked copies both ways, chunk %8d: %.1f GB/s each way" % (chunk, raw_chunked(chunk)))
    ctx = g.Context(0)
    ctx.set_lut_from_cube(g.parse_cube(frames.cube_text_3d(65)))
    src = frames.frame_grad(w, h).reshape(-1)
    for mode in (0, 3, 1, 2):
      ctx.set_option("host.dbg_mode", mode)
      print("host.dbg_mode", mode, "(0 normal, 3 one-row kernel, 1 no kernel, 2 no kernel and no cross-stream waits)")
      for nb in (8,):
            hin = [torch.from_numpy(src.copy()).pin_memory() for _ in range(nb)]
            hout = [torch.empty_like(t).pin_memory() for t in hin]
            fi = frame_array([frame_of(t, w, h, "RGBA") for t in hin])
            fo = frame_array([frame_of(t, w, h, "RGBA") for t in hout])
            for chunk in (2 << 20, 4 << 20, 16 << 20):
                ctx.set_option("host.chunk_bytes", chunk)
                for _ in range(2):
                    ctx.colorlut_batch(fi, fo)
                c0, w0, t0c = (ctx.get_option("host.dbg_chunks"), ctx.get_option("host.dbg_wait_ns"),
                               ctx.get_option("host.dbg_call_ns"))
                iters = max(2, 64 // nb)
                t0 = time.perf_counter()
                for _ in range(iters):
                    ctx.colorlut_batch(fi, fo)
                dt = time.perf_counter() - t0
                chunks = ctx.get_option("host.dbg_chunks") - c0
                wait = ctx.get_option("host.dbg_wait_ns") - w0
                call = ctx.get_option("host.dbg_call_ns") - t0c
                print("frames/call %2d chunk %8d: %6.0f frames/s  %5.1f GB/s  chunks %4d  per chunk: total %6.1f us, "
                      "blocked %6.1f us, enqueue %6.1f us" %
                      (nb, chunk, nb * iters / dt, nb * iters * fb / dt / 1e9, chunks, call / chunks / 1e3,
                       wait / chunks / 1e3, (call - wait) / chunks / 1e3))


if __name__ == "__main__":
    main()
