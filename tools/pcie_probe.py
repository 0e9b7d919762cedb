"""Pinned-memory PCIe peaks on this box (H2D, D2H, both at once) and the e2e host path at several
chunk sizes — the denominator for bench.py's e2e number (development aid)."""
import sys
import time

import torch

sys.path.insert(0, ".")
import gst_plugins_rs_b200 as g
from gst_plugins_rs_b200 import frames
from gst_plugins_rs_b200.api import frame_array, frame_of


def bw(fn, nbytes, iters=20):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    torch.cuda.synchronize()
    return nbytes * iters / (time.perf_counter() - t0) / 1e9


def main():
    n = 256 << 20
    h1 = torch.empty(n, dtype=torch.uint8).pin_memory()
    h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
    d1 = torch.empty(n, dtype=torch.uint8, device="cuda")
    d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    print("H2D  %.1f GB/s" % bw(lambda: d1.copy_(h1, non_blocking=True), n))
    print("D2H  %.1f GB/s" % bw(lambda: h2.copy_(d2, non_blocking=True), n))

    def both():
        with torch.cuda.stream(s1):
            d1.copy_(h1, non_blocking=True)
        with torch.cuda.stream(s2):
            h2.copy_(d2, non_blocking=True)
    print("both %.1f GB/s each way" % bw(both, n))

    w, h, nb = 3840, 2160, 8
    ctx = g.Context(0)
    ctx.set_lut_from_cube(g.parse_cube(frames.cube_text_3d(65)))
    src = frames.frame_grad(w, h).reshape(-1)
    hin = [torch.from_numpy(src.copy()).pin_memory() for _ in range(nb)]
    hout = [torch.empty_like(t).pin_memory() for t in hin]
    fi = frame_array([frame_of(t, w, h, "RGBA") for t in hin])
    fo = frame_array([frame_of(t, w, h, "RGBA") for t in hout])
    for slots in (3, 4, 6, 8):
        ctx.set_option("host.slots", slots)
        for chunk in (1 << 20, 2 << 20, 4 << 20, 8 << 20, 16 << 20, 33177600):
            ctx.set_option("host.chunk_bytes", chunk)
            gb = bw(lambda: ctx.colorlut_batch(fi, fo), nb * w * h * 4, iters=8)
            print("e2e colorlut, 8 frames per call, slots %d chunk %8d B: %.1f GB/s each way = %.0f frames/s" %
                  (slots, chunk, gb, gb * 1e9 / (w * h * 4)))
    # single-frame latency (one call per frame, like the element)
    f1, o1 = frame_array([frame_of(hin[0], w, h, "RGBA")]), frame_array([frame_of(hout[0], w, h, "RGBA")])
    for slots in (3, 4, 6, 8):
        ctx.set_option("host.slots", slots)
        for chunk in (1 << 20, 2 << 20, 4 << 20, 8 << 20):
            ctx.set_option("host.chunk_bytes", chunk)
            gb = bw(lambda: ctx.colorlut_batch(f1, o1), w * h * 4, iters=30)
            print("e2e single-frame calls, slots %d chunk %d: %.0f frames/s" % (slots, chunk, gb * 1e9 / (w * h * 4)))
    # pageable frames (numpy): bounce path vs page-locked in place on second sight
    import numpy as np
    pin, pout = src.copy(), np.zeros_like(src)
    fp, op = frame_array([frame_of(pin, w, h, "RGBA")]), frame_array([frame_of(pout, w, h, "RGBA")])
    ctx.set_option("host.slots", 4)
    ctx.set_option("host.chunk_bytes", 0)
    for reg in (0, 1):
        ctx.set_option("host.register", reg)
        for _ in range(3):  # second sight registers (a few ms, once)
            ctx.colorlut_batch(fp, op)
        gb = bw(lambda: ctx.colorlut_batch(fp, op), w * h * 4, iters=30)
        print("e2e pageable single-frame calls, host.register=%d: %.0f frames/s" % (reg, gb * 1e9 / (w * h * 4)))


if __name__ == "__main__":
    main()
