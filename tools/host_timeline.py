import sys, torch
sys.path.insert(0, ".")
import gst_plugins_rs_b200 as g
from gst_plugins_rs_b200 import frames
from gst_plugins_rs_b200.api import frame_array, frame_of
w, h, nb = 3840, 2160, 4
ctx = g.Context(0)
ctx.set_lut_from_cube(g.parse_cube(frames.cube_text_3d(65)))
src = frames.frame_grad(w, h).reshape(-1)
hin = [torch.from_numpy(src.copy()).pin_memory() for _ in range(nb)]
hout = [torch.empty_like(t).pin_memory() for t in hin]
fi = frame_array([frame_of(t, w, h, "RGBA") for t in hin]); fo = frame_array([frame_of(t, w, h, "RGBA") for t in hout])
ctx.set_option("host.chunk_bytes", 16 << 20)
for _ in range(3): ctx.colorlut_batch(fi, fo)
ctx.set_option("host.dbg_mode", 4)
ctx.colorlut_batch(fi, fo)
