"""Per-call cost of single-frame device-memory calls (what a memory:CUDAMemory element issues): host time
per call and frames/s, auto policy vs pinned kernel (dev aid)."""
import sys
import time
sys.path.insert(0, ".")
import torch
import gst_plugins_rs_b200 as g
from gst_plugins_rs_b200 import frames
from gst_plugins_rs_b200.api import frame_of
ctx = g.Context(0)
ctx.set_lut_from_cube(g.parse_cube(frames.cube_text_3d(33)))
p = g.HsvFilterParams(37.5, 1.2, 0.05, 0.9, 0.02)
for (w, h) in ((1920, 1080), (3840, 2160)):
    ring = 24
    src = [torch.from_numpy(frames.frame_noise(w, h, i % 3).reshape(-1).copy()).cuda() for i in range(ring)]
    dst = [torch.empty_like(s) for s in src]
    fi = [frame_of(s, w, h, "RGBA") for s in src]
    fo = [frame_of(d, w, h, "RGBA") for d in dst]
    for name, opt, val, call in (("colorlut auto", "lut.path", 0, lambda k: ctx.colorlut(fi[k], fo[k])),
                                 ("colorlut lut.path=4", "lut.path", 4, lambda k: ctx.colorlut(fi[k], fo[k])),
                                 ("hsvfilter auto", "hsv.path", 0, lambda k: ctx.hsvfilter(fo[k], p)),
                                 ("hsvfilter hsv.path=2", "hsv.path", 2, lambda k: ctx.hsvfilter(fo[k], p))):
        ctx.set_option(opt, val)
        for k in range(60):
            call(k % ring)
        ctx.synchronize()
        n = 3000
        t0 = time.perf_counter()
        for k in range(n):
            call(k % ring)
        t_host = time.perf_counter() - t0
        ctx.synchronize()
        t_all = time.perf_counter() - t0
        print("%dx%d %-22s host %.2f us/call, %.0f frames/s (%.1f%% of the HBM peak)" %
              (w, h, name, t_host / n * 1e6, n / t_all, 8 * w * h * n / t_all / 1e9 / 6548.5 * 100))
