"""Zero-copy experiment: launch the kernels directly on page-locked host frames (UVA-mapped) instead
of staging them through device buffers with the copy engines (dev aid)."""
import sys
import time
import numpy as np
import torch
sys.path.insert(0, ".")
import gst_plugins_rs_b200 as g
from gst_plugins_rs_b200 import frames
from gst_plugins_rs_b200.api import frame_array, frame_of
import oracle

w, h = 3840, 2160
fb = w * h * 4
ctx = g.Context(0)
text = frames.cube_text_3d(33)
ctx.set_lut_from_cube(g.parse_cube(text))
src = frames.frame_noise(w, h, 1).reshape(-1)
for nb in (1, 8):
    hin = [torch.from_numpy(src.copy()).pin_memory() for _ in range(nb)]
    hout = [torch.zeros_like(t).pin_memory() for t in hin]
    for label, mem in (("staged (copy engines)", 0), ("zero-copy (kernel on host pointers)", 1)):
        fi, fo = [frame_of(t, w, h, "RGBA") for t in hin], [frame_of(t, w, h, "RGBA") for t in hout]
        for f in fi + fo:
            f.memory = mem
        fi, fo = frame_array(fi), frame_array(fo)
        for _ in range(3):
            ctx.colorlut_batch(fi, fo)
            ctx.synchronize()
        iters = 40 // nb + 2
        t0 = time.perf_counter()
        for _ in range(iters):
            ctx.colorlut_batch(fi, fo)
            ctx.synchronize()
        dt = time.perf_counter() - t0
        ok = np.array_equal(hout[0].numpy()[: w * 64 * 4], oracle.colorlut(oracle.Lut(text=text), src[: w * 64 * 4], w, 64))
        print("%d frame(s) per call, %-38s %6.0f frames/s  %5.1f GB/s each way  parity %s" %
              (nb, label, nb * iters / dt, nb * iters * fb / dt / 1e9, "ok" if ok else "MISMATCH"))
