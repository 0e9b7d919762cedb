import sys
sys.path.insert(0, ".")
import torch
import gst_plugins_rs_b200 as g
from gst_plugins_rs_b200 import frames
from gst_plugins_rs_b200.api import frame_array, frame_of
w, h, nb = 3840, 2160, 16
ctx = g.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
for interp in (0, 1):
    ctx.set_option("lut.interpolation", interp)
    for content in ("noise", "rand"):
        ctx.set_lut_from_cube(g.parse_cube(frames.cube_text_3d(65)))
        base = [torch.from_numpy(frames.frame_of_class(content, w, h, i % 4).reshape(-1).copy()).cuda() for i in range(nb)]
        dst = [torch.empty_like(b) for b in base]
        fin = frame_array([frame_of(b, w, h, "RGBA") for b in base])
        fout = frame_array([frame_of(d, w, h, "RGBA") for d in dst])
        row = []
        for i in range(24):
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            ctx.colorlut_batch(fin, fout)
            e1.record()
            if i % 4 == 3 or i < 4:
                torch.cuda.synchronize()
                row.append("%d:%.2f" % (ctx.get_option("lut.path_active"), e0.elapsed_time(e1)))
            else:
                row.append("%d" % ctx.get_option("lut.path_active"))
        torch.cuda.synchronize()
        print("interp", interp, content, " ".join(row))
