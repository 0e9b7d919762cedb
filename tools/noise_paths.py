import sys
sys.path.insert(0, ".")
import torch
import gst_plugins_rs_b200 as g
from gst_plugins_rs_b200 import frames
from gst_plugins_rs_b200.api import frame_array, frame_of
w, h, nb = 3840, 2160, 16
ctx = g.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
def timed(fn, iters=20):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
for amp in (1, 2, 4, 8):
    base = [torch.from_numpy(frames.frame_noise(w, h, i, amp).reshape(-1).copy()).cuda() for i in range(nb)]
    dst = [torch.empty_like(b) for b in base]
    fin = frame_array([frame_of(b, w, h, "RGBA") for b in base])
    fout = frame_array([frame_of(d, w, h, "RGBA") for d in dst])
    for n in (33, 65):
        ctx.set_lut_from_cube(g.parse_cube(frames.cube_text_3d(n)))
        row = []
        for path in (1, 2, 3, 4):
            ctx.set_option("lut.path", path)
            ms = timed(lambda: ctx.colorlut_batch(fin, fout))
            row.append(8 * w * h * nb / ms / 1e6 / 6548.5 * 100)
        print(f"noise ±{amp} N={n}: direct {row[0]:5.1f}%  R {row[1]:5.1f}%  RG {row[2]:5.1f}%  baked {row[3]:5.1f}%")
