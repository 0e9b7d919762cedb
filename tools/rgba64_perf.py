"""RGBA64 colorlut timing across content classes and LUT sizes (dev aid; B200VF_LIB selects a build)."""
import sys
sys.path.insert(0, ".")
import numpy as np
import torch
import gst_plugins_rs_b200 as g
from gst_plugins_rs_b200 import frames
from gst_plugins_rs_b200.api import frame_array, frame_of
import oracle

w, h, nb = 3840, 2160, 8
ctx = g.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for n in (33, 65, 17, 20):
    text = frames.cube_text_3d(n)
    ctx.set_lut_from_cube(g.parse_cube(text))
    row = []
    for content in ("bars", "grad", "noise", "rand"):
        src = [(frames.frame_of_class(content, w, h, i).reshape(-1).astype("<u2") * 257).view(np.uint8) for i in range(2)]
        base = [torch.from_numpy(src[i % 2]).cuda().clone() for i in range(nb)]
        dst = [torch.empty_like(b) for b in base]
        fin = frame_array([frame_of(b, w, h, "RGBA64_LE") for b in base])
        fout = frame_array([frame_of(d, w, h, "RGBA64_LE") for d in dst])
        ms = timed(lambda: ctx.colorlut_batch(fin, fout))
        row.append("%s %5.1f%%" % (content, 16 * w * h * nb / ms / 1e6 / 6548.5 * 100))
        if content == "noise" and n in (33, 20):  # parity spot check on a crop
            hh = 64
            want = oracle.colorlut(oracle.Lut(text=text), src[0][: w * hh * 8], w, hh, "RGBA64_LE")
            got = dst[0].cpu().numpy()[: w * hh * 8]
            row.append("parity %s" % ("ok" if np.array_equal(got, want) else "MISMATCH"))
    print(f"N={n:3d} built={ctx.get_option('lut.tables_built')}  " + "  ".join(row))
