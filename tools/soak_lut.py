"""One-off soak (dev aid): random LUT sizes / values / domains x random frames, every LUT path,
8- and 16-bit, against the oracle.  usage: python tools/soak_lut.py [n_luts]"""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import gst_plugins_rs_b200 as g, oracle, util
from gst_plugins_rs_b200 import frames

N = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rng = np.random.default_rng(4242)
ctx = g.Context(0)
W, H = 384, 64
bad = 0
for i in range(N):
    kind3 = rng.random() < 0.75
    n = int(rng.integers(2, 41)) if kind3 else int(rng.choice([2, 3, 17, 256, 1000, 4096]))
    lo, hi = (-0.3, 1.3) if rng.random() < 0.4 else (0.0, 1.0)
    cnt = n ** 3 if kind3 else n
    vals = rng.uniform(lo, hi, size=(cnt, 3))
    dom = None
    if rng.random() < 0.5:
        mn = rng.uniform(-1, 0.5, 3); mx = mn + rng.uniform(0.1, 3, 3)
        dom = (tuple(mn), tuple(mx))
    text = (frames.cube_text_3d(n, vals, *(dom or (None, None))) if kind3
            else frames.cube_text_1d(n, vals, *(dom or (None, None))))
    lut = oracle.Lut(text=text)
    ctx.set_lut_from_cube(g.parse_cube(text))
    src8 = frames.random_bytes(W * H * 4, i)
    if i % 3 == 0:
        src8 = frames.frame_noise(W, H, i, 3).reshape(-1)
    want8 = oracle.colorlut(lut, src8, W, H)
    for path in ((0, 1, 2, 3, 4) if kind3 else (0,)):
        for math in (0, 1):
            ctx.set_option("lut.path", path); ctx.set_option("hsv.math", math)
            got = util.gpu_colorlut(ctx, src8, W, H)
            if not np.array_equal(got, want8):
                bad += 1; print("MISMATCH 8-bit", i, n, kind3, dom, path, math, int((got != want8).sum()))
    ctx.set_option("lut.path", 0); ctx.set_option("hsv.math", 0)
    src16 = frames.random_bytes(W * H * 8, 1000 + i)
    for fmt in ("RGBA64_LE", "RGBA64_BE"):
        got = util.gpu_colorlut(ctx, src16, W, H, fmt)
        want = oracle.colorlut(lut, src16, W, H, fmt)
        if not np.array_equal(got, want):
            bad += 1; print("MISMATCH", fmt, i, n, kind3, dom, int((got != want).sum()))
print("colorlut soak:", N, "LUTs,", bad, "mismatching runs")
