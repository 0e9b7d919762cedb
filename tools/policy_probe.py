import sys
sys.path.insert(0, ".")
import torch
import gst_plugins_rs_b200 as g
from gst_plugins_rs_b200 import frames
from gst_plugins_rs_b200.api import frame_array, frame_of
w, h, nb = 3840, 2160, 16
ctx = g.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
for content in ("rand", "noise"):
    base = [torch.from_numpy(frames.frame_of_class(content, w, h, i).reshape(-1).copy()).cuda() for i in range(nb)]
    dst = [torch.empty_like(b) for b in base]
    fin = frame_array([frame_of(b, w, h, "BGRx") for b in base])
    fout = frame_array([frame_of(d, w, h, "RGBA") for d in dst])
    dp = g.HsvDetectorParams(120, 30, 0.6, 0.4, 0.6, 0.4 + (0.01 if content == "noise" else 0))
    for sync in (True, False):
        dp = g.HsvDetectorParams(120, 30, 0.6, 0.4, 0.6, dp.value_var + 0.001)
        row = []
        for i in range(30):
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            ctx.hsvdetector_batch(fin, fout, dp)
            e1.record()
            if sync:
                torch.cuda.synchronize()
                row.append("%s%.2f" % ("T" if ctx.get_option("hsv.table_active") else "c", e0.elapsed_time(e1)))
            else:
                row.append("T" if ctx.get_option("hsv.table_active") else "c")
        torch.cuda.synchronize()
        print(content, "sync" if sync else "async", " ".join(row))
