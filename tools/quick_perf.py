"""Quick device-resident timing of the three elements (development aid, not the bench)."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import gst_plugins_rs_b200 as g
from gst_plugins_rs_b200 import frames
from gst_plugins_rs_b200.api import frame_array, frame_of

PEAK = 6548.5


def timed(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    w, h, nb = 3840, 2160, 16  # 16 frames x 33 MB x2 = 1 GB working set > L2
    ctx = g.Context(0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    out = {}
    for content in ("bars", "grad", "rand"):
        base = [torch.from_numpy(frames.frame_of_class(content, w, h, i).reshape(-1).copy()).cuda()
                for i in range(nb)]
        dst = [torch.empty_like(b) for b in base]
        fin = frame_array([frame_of(b, w, h, "RGBA") for b in base])
        fout = frame_array([frame_of(d, w, h, "RGBA") for d in dst])
        bytes_per = 8 * w * h * nb
        # hsvfilter (in place on dst copies so content stays as generated for each timing)
        p = g.HsvFilterParams(37.5, 1.2, 0.05, 0.9, 0.02)
        for d, b in zip(dst, base):
            d.copy_(b)
        for math in (0, 1):
            ctx.set_option("hsv.math", math)
            ms = timed(lambda: ctx.hsvfilter_batch(fout, p))
            out[f"hsvfilter/{content}/math{math}"] = (ms, bytes_per / ms / 1e6)
        ctx.set_option("hsv.math", 0)
        dp = g.HsvDetectorParams(120, 30, 0.6, 0.4, 0.6, 0.4)
        fin_b = frame_array([frame_of(b, w, h, "BGRx") for b in base])
        ms = timed(lambda: ctx.hsvdetector_batch(fin_b, fout, dp))
        out[f"hsvdetector/{content}"] = (ms, bytes_per / ms / 1e6)
        for n in (33, 65):
            ctx.set_lut_from_cube(g.parse_cube(frames.cube_text_3d(n)))
            for path in (1, 2, 3):
                ctx.set_option("lut.path", path)
                ms = timed(lambda: ctx.colorlut_batch(fin, fout))
                out[f"colorlut{n}/{content}/path{path}"] = (ms, bytes_per / ms / 1e6)
                ms = timed(lambda: ctx.chain_lut_hsv_batch(fin, fout, p))
                out[f"chain{n}/{content}/path{path}"] = (ms, bytes_per / ms / 1e6)
        # plain copy for reference
        ms = timed(lambda: [d.copy_(b) for d, b in zip(dst, base)])
        out[f"copy/{content}"] = (ms, bytes_per / ms / 1e6)
    for k, (ms, gbs) in out.items():
        print(f"{k:36s} {ms:9.3f} ms/batch  {gbs:8.1f} GB/s  {gbs / PEAK * 100:5.1f}% of measured HBM"
              f"  {nb / ms * 1e3:9.0f} frames/s")
    json.dump(out, open("gpurun_out/quick_perf.json", "w"), indent=1)


if __name__ == "__main__":
    main()
