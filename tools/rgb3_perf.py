"""3-byte formats through the function tables: hsvfilter RGB (6 B/px) and hsvdetector RGB -> RGBA
(7 B/px) on grad / noise, device-resident (dev aid; B200VF_LIB selects a build)."""
import sys
sys.path.insert(0, ".")
import numpy as np
import torch
import gst_plugins_rs_b200 as g
from gst_plugins_rs_b200 import frames
from gst_plugins_rs_b200.api import frame_array, frame_of

w, h, nb = 3840, 2160, 16
ctx = g.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
ctx.set_option("hsv.path", 2)


def timed(fn, iters=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


p = g.HsvFilterParams(37.5, 1.2, 0.05, 0.9, 0.02)
dp = g.HsvDetectorParams(120, 30, 0.6, 0.4, 0.6, 0.4)
out = []
for content in ("grad", "noise"):
    rgb = [np.ascontiguousarray(frames.frame_of_class(content, w, h, i).reshape(h, w, 4)[:, :, :3]).reshape(-1) for i in range(2)]
    src = [torch.from_numpy(rgb[i % 2]).cuda().clone() for i in range(nb)]
    dst3 = [torch.empty_like(s) for s in src]
    dst4 = [torch.empty(w * h * 4, dtype=torch.uint8, device="cuda") for _ in src]
    f3 = frame_array([frame_of(t, w, h, "RGB") for t in src])
    o4 = frame_array([frame_of(t, w, h, "RGBA") for t in dst4])
    # hsvfilter works in place: filter copies so that the content class survives the repeats
    def filt():
        for d, s in zip(dst3, src):
            d.copy_(s)
        ctx.hsvfilter_batch(frame_array([frame_of(t, w, h, "RGB") for t in dst3]), p)
    def copy_only():
        for d, s in zip(dst3, src):
            d.copy_(s)
    t_f = timed(filt) - timed(copy_only)
    t_d = timed(lambda: ctx.hsvdetector_batch(f3, o4, dp))
    out.append("%s: hsvfilter RGB %5.1f %%  hsvdetector RGB->RGBA %5.1f %%" %
               (content, 6 * w * h * nb / t_f / 1e6 / 6548.5 * 100, 7 * w * h * nb / t_d / 1e6 / 6548.5 * 100))
print("  |  ".join(out))
