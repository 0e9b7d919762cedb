"""hsvfilter / hsvdetector with hundreds of random settings — uniform ranges, extreme magnitudes,
arbitrary float bit patterns (denormals, huge, NaN, inf) — on a frame holding random pixels, all
greys and all primaries' ramps.  The HSV→RGB half of the fast path cannot be enumerated (its input
is continuous), so this is where its shortcuts meet arbitrary hue/saturation/value floats."""
import struct

import numpy as np
import pytest

import util
from gst_plugins_rs_b200 import frames

pytestmark = pytest.mark.gpu
W, H = 512, 256


def _frame():
    px = frames.frame_rand(W, H, 4, 99).reshape(-1, 4)
    ramp = np.arange(256, dtype=np.uint8)
    px[:256, :3] = ramp[:, None]
    for c in range(3):
        px[256 * (c + 1):256 * (c + 2), :3] = 0
        px[256 * (c + 1):256 * (c + 2), c] = ramp
    px[1024:1280, 0], px[1024:1280, 1], px[1024:1280, 2] = 255, ramp, 255 - ramp
    return px.reshape(-1)


def _random_float(rng, kind):
    if kind == 0:
        return float(rng.uniform(-400, 400))
    if kind == 1:
        return float(rng.uniform(-2, 3))
    if kind == 2:
        return float(np.float32(10.0) ** np.float32(rng.uniform(-40, 38)) * rng.choice([-1, 1]))
    if kind == 3:
        return struct.unpack("<f", struct.pack("<I", int(rng.integers(0, 1 << 32))))[0]
    return float(rng.choice([0.0, -0.0, 1.0, -1.0, 360.0, -360.0, 180.0, 720.0, 59.999996, 60.0,
                             1e-45, 3.4028235e38, float("inf"), float("-inf"), float("nan")]))


def test_hsvfilter_random_settings(ctx, orc):
    rng = np.random.default_rng(2026)
    src = _frame()
    for i in range(160):
        kinds = rng.integers(0, 5, size=5) if i % 3 else [0, 1, 1, 1, 1]
        s = tuple(_random_float(rng, int(k)) for k in kinds)
        got = util.gpu_hsvfilter(ctx, src, W, H, "RGBA", s)
        want = orc.hsvfilter(src, W, H, "RGBA", s)
        assert np.array_equal(got, want), f"settings {s!r}: {int((got != want).sum())} bytes differ"


def test_hsvdetector_random_settings(ctx, orc):
    rng = np.random.default_rng(17)
    src = _frame()
    hits = 0
    for i in range(160):
        if i % 3:
            kinds = rng.integers(0, 5, size=6)
            s = tuple(_random_float(rng, int(k)) for k in kinds)
        else:  # in-range properties (hsvdetector/imp.rs:164-214)
            s = (float(rng.uniform(-720, 720)), float(rng.uniform(0, 180)), float(rng.uniform(0, 1)),
                 float(rng.uniform(0, 1)), float(rng.uniform(0, 1)), float(rng.uniform(0, 1)))
        got = util.gpu_hsvdetector(ctx, src, W, H, "RGBx", "BGRA", s)
        want = orc.hsvdetector(src, W, H, "RGBx", "BGRA", s)
        assert np.array_equal(got, want), f"settings {s!r}: {int((got != want).sum())} bytes differ"
        hits += int((want.reshape(-1, 4)[:, 3] == 255).any())
    assert hits > 40  # the sweep does exercise matching windows
