"""A/B timing of library variants (variants/lib_*.so) on the three 4K workloads (dev aid).
usage: python tools/ab_perf.py [lib ...]   (no args = default lib + every variants/lib_*.so)"""
import glob
import json
import os
import subprocess
import sys

CHILD = r"""
import sys, json
sys.path.insert(0, ".")
import torch
import gst_plugins_rs_b200 as g
from gst_plugins_rs_b200 import frames
from gst_plugins_rs_b200.api import frame_array, frame_of
w, h, nb = 3840, 2160, 16
ctx = g.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
def timed(fn, iters=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
out = {}
for content in ("bars", "grad", "noise", "rand"):
    base = [torch.from_numpy(frames.frame_of_class(content, w, h, i).reshape(-1).copy()).cuda() for i in range(nb)]
    dst = [torch.empty_like(b) for b in base]
    fin = frame_array([frame_of(b, w, h, "RGBA") for b in base])
    fout = frame_array([frame_of(d, w, h, "RGBA") for d in dst])
    finb = frame_array([frame_of(b, w, h, "BGRx") for b in base])
    for n in (33, 65):
        ctx.set_lut_from_cube(g.parse_cube(frames.cube_text_3d(n)))
        out[f"lut{n}/{content}"] = timed(lambda: ctx.colorlut_batch(fin, fout))
        if n == 33:
            ctx.set_option("lut.path", 4)
            out[f"baked/{content}"] = timed(lambda: ctx.colorlut_batch(fin, fout))
            ctx.set_option("lut.path", 0)
    if content == "grad":
        p = g.HsvFilterParams(37.5, 1.2, 0.05, 0.9, 0.02)
        out["hsvfilter"] = timed(lambda: ctx.hsvfilter_batch(fout, p))
        dp = g.HsvDetectorParams(120, 30, 0.6, 0.4, 0.6, 0.4)
        out["hsvdetector"] = timed(lambda: ctx.hsvdetector_batch(finb, fout, dp))
        out["chain33"] = timed(lambda: ctx.chain_lut_hsv_batch(fin, fout, p))
print(json.dumps(out))
"""


def main():
    libs = sys.argv[1:] or ([""] + sorted(glob.glob("variants/lib_*.so")))
    bytes_per = 8 * 3840 * 2160 * 16
    rows = {}
    for lib in libs:
        env = dict(os.environ)
        if lib:
            env["B200VF_LIB"] = os.path.abspath(lib)
        r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
        if r.returncode != 0:
            print(lib, "FAILED", r.stderr[-500:])
            continue
        rows[lib or "default"] = json.loads(r.stdout.strip().splitlines()[-1])
    keys = list(next(iter(rows.values())))
    print("%-22s" % "variant" + "".join("%13s" % k for k in keys))
    for lib, d in rows.items():
        print("%-22s" % os.path.basename(lib) +
              "".join("%12.1f%%" % (bytes_per / d[k] / 1e6 / 6548.5 * 100) for k in keys))


if __name__ == "__main__":
    main()
