"""Synchronous host-frame calls against "host.async" = 1 with one call of latency (wait for call
k-1 after queueing call k), 1 and 8 pinned 4K frames per call, several chunk sizes, next to the
box's bidirectional PCIe peak (development aid; results in profiles/)."""
import sys
import time

import torch

sys.path.insert(0, ".")
import gst_plugins_rs_b200 as g
from gst_plugins_rs_b200 import frames
from gst_plugins_rs_b200.api import frame_array, frame_of

W, H = 3840, 2160


def peak():
    n = 256 << 20
    h1 = torch.empty(n, dtype=torch.uint8).pin_memory()
    h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
    d1 = torch.empty(n, dtype=torch.uint8, device="cuda")
    d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def both():
        with torch.cuda.stream(s1):
            d1.copy_(h1, non_blocking=True)
        with torch.cuda.stream(s2):
            h2.copy_(d2, non_blocking=True)
    both()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        both()
    torch.cuda.synchronize()
    return n * 10 / (time.perf_counter() - t0) / 1e9


def main():
    pk = peak()
    print("bidirectional pinned-copy peak: %.1f GB/s each way = %.0f 4K RGBA frames/s" % (pk, pk * 1e9 / (W * H * 4)))
    ctx = g.Context(0)
    ctx.set_lut_from_cube(g.parse_cube(frames.cube_text_3d(65)))
    src = frames.frame_grad(W, H).reshape(-1)
    n_buf = 16
    hin = [torch.from_numpy(src.copy()).pin_memory() for _ in range(n_buf)]
    hout = [torch.empty_like(t).pin_memory() for t in hin]
    for per_call in (1, 8):
        sets = []
        for k in range(n_buf // per_call):
            a = frame_array([frame_of(t, W, H, "RGBA") for t in hin[k * per_call:(k + 1) * per_call]])
            b = frame_array([frame_of(t, W, H, "RGBA") for t in hout[k * per_call:(k + 1) * per_call]])
            sets.append((a, b))
        calls = 240 // per_call
        for mode in (0, 1):
            for chunk in (0, 4 << 20, 8 << 20, 17 << 20, W * H * 4):
                ctx.set_option("host.async", mode)
                ctx.set_option("host.chunk_bytes", chunk)
                for rep in range(2):  # first pass warms up
                    ctx.synchronize()
                    t0 = time.perf_counter()
                    prev = None
                    for c in range(calls):
                        a, b = sets[c % len(sets)]
                        ctx.colorlut_batch(a, b)
                        if mode:
                            t = ctx.host_ticket()
                            if prev is not None:
                                ctx.host_wait(prev)
                            prev = t
                    ctx.synchronize()
                    dt = time.perf_counter() - t0
                fps = calls * per_call / dt
                print("%d frame(s)/call  %-5s chunk %9d B: %7.0f frames/s = %.1f GB/s each way = %.0f %% of the peak" %
                      (per_call, "async" if mode else "sync", chunk, fps, fps * W * H * 4 / 1e9,
                       100 * fps * W * H * 4 / 1e9 / pk))
    ctx.set_option("host.async", 0)
    ctx.close()


if __name__ == "__main__":
    main()
