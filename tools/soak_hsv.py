import sys, struct
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
import gst_plugins_rs_b200 as g, oracle, util
from gst_plugins_rs_b200 import frames
from gst_plugins_rs_b200.api import frame_of
import test_gpu_hsv_random_settings as T
W, H = 256, 96
px = frames.frame_rand(W, H, 4, 5).reshape(-1, 4)
ramp = np.arange(256, dtype=np.uint8)
px[:256, :3] = ramp[:, None]
for c in range(3):
    px[256*(c+1):256*(c+2), :3] = 0; px[256*(c+1):256*(c+2), c] = ramp
src = px.reshape(-1)
ctx = g.Context(0)
# optional second argument: "hsv.path" (2 = every settings tuple goes through a freshly built table)
ctx.set_option("hsv.path", int(sys.argv[2]) if len(sys.argv) > 2 else 0)
rng = np.random.default_rng(12345)
bad = 0
N = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
for i in range(N):
    kinds = rng.integers(0, 5, size=5)
    if i % 2 == 0: kinds = [rng.integers(0, 5), 1, 1, 1, 1]
    s = tuple(T._random_float(rng, int(k)) for k in kinds)
    t = torch.from_numpy(src.copy()).cuda()
    ctx.hsvfilter(frame_of(t, W, H, "RGBA"), g.HsvFilterParams(*s)); ctx.synchronize()
    got = t.cpu().numpy()
    want = oracle.hsvfilter(src, W, H, "RGBA", s)
    if not np.array_equal(got, want):
        bad += 1
        print("MISMATCH", s, int((got != want).sum()))
print("hsvfilter soak:", N, "settings,", bad, "mismatching")
bad = 0
for i in range(N):
    s = tuple(T._random_float(rng, int(k)) for k in rng.integers(0, 5, size=6))
    if i % 2 == 0:
        s = (float(rng.uniform(-720, 720)), float(rng.uniform(0, 180)), float(rng.uniform(0, 1)), float(rng.uniform(0, 1)), float(rng.uniform(0, 1)), float(rng.uniform(0, 1)))
    got = util.gpu_hsvdetector(ctx, src, W, H, "BGRx", "RGBA", s)
    want = oracle.hsvdetector(src, W, H, "BGRx", "RGBA", s)
    if not np.array_equal(got, want):
        bad += 1; print("MISMATCH det", s, int((got != want).sum()))
print("hsvdetector soak:", N, "settings,", bad, "mismatching")
