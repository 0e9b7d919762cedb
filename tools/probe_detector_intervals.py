"""Round-2 de-risking (CPU only): for the hsvdetector saturation & value windows, is the set of
passing `min` bytes an interval for every `max` byte, for arbitrary settings?  Brute force in f32."""
import numpy as np
F = np.float32
mx = np.arange(256, dtype=np.int64)[:, None]
mn = np.arange(256, dtype=np.int64)[None, :]
valid = mn <= mx
value = (mx.astype(F) / F(255))
chroma = (value - (mn.astype(F) / F(255))).astype(F)
with np.errstate(divide="ignore", invalid="ignore"):
    sat = np.where(value == 0, F(0), (chroma / value).astype(F)).astype(F)
sat = np.clip(sat, F(0), F(1))          # rs_clamp; no NaN possible here
val = np.clip(value, F(0), F(1)) + np.zeros_like(sat)
rng = np.random.default_rng(1)
def weird():
    r = rng.integers(0, 4)
    if r == 0: return F(rng.uniform(0, 1))
    if r == 1: return F(rng.choice([0, 1, 0.5, -0.0, 2, -1, np.inf, -np.inf, np.nan, 1e-8, 0.15, 0.3]))
    if r == 2: return np.frombuffer(rng.bytes(4), dtype=F)[0]
    return F(rng.uniform(-0.5, 1.5))
bad = 0
N = 20000
for it in range(N):
    sr, sv, vr, vv = weird(), weird(), weird(), weird()
    with np.errstate(invalid="ignore", over="ignore"):
        ok = (np.abs((sat - sr).astype(F)) <= sv) & (np.abs((val - vr).astype(F)) <= vv) & valid
    # interval check per row: the passing mins must be contiguous
    for m in range(256):
        idx = np.nonzero(ok[m, :m + 1])[0]
        if idx.size and idx[-1] - idx[0] + 1 != idx.size:
            bad += 1
            print("NOT AN INTERVAL", sr, sv, vr, vv, m, idx[:10]); break
print(N, "settings;", bad, "with a non-interval row")
