"""The C oracle and the independent numpy-f32 restatement (np_emulation.py) must agree bit for
bit on seeded inputs: the strongest pin available for the pixel loops, which no reference
test covers."""
import numpy as np
import pytest

import np_emulation as npe
from gst_plugins_rs_b200 import frames

SETTINGS = [(0.0, 1.0, 0.0, 1.0, 0.0), (37.5, 1.2, 0.05, 0.9, 0.02), (-123.25, 0.7, -0.1, 1.3, 0.1),
            (1234.5, 1.0, 0.25, 1.0, -0.25), (float("nan"), float("nan"), 0.0, 1.0, 0.0)]


def _px(n, seed):
    return frames.random_bytes(n * 4, seed).reshape(n, 4)


@pytest.mark.parametrize("settings", SETTINGS)
def test_hsvfilter_oracle_equals_numpy(orc, settings):
    px = _px(1 << 20, 1)
    # include every grey and every primary ramp
    px[:256, :3] = np.arange(256, dtype=np.uint8)[:, None]
    want = px.copy()
    want[:, 0], want[:, 1], want[:, 2] = npe.hsvfilter_rgb(px[:, 0], px[:, 1], px[:, 2], settings)
    got = orc.hsvfilter(px, 1 << 10, 1 << 10, "RGBA", settings).reshape(-1, 4)
    assert np.array_equal(got, want)
    # BGR family reads/writes reversed (hsvutils.rs:88-128, 167-198)
    want_bgr = px.copy()
    r, g, b = npe.hsvfilter_rgb(px[:, 3], px[:, 2], px[:, 1], settings)
    want_bgr[:, 3], want_bgr[:, 2], want_bgr[:, 1] = r, g, b
    got = orc.hsvfilter(px, 1 << 10, 1 << 10, "xBGR", settings).reshape(-1, 4)
    assert np.array_equal(got, want_bgr)


@pytest.mark.parametrize("settings", [(0.0, 10.0, 0.0, 0.15, 0.0, 0.3), (120.0, 30.0, 0.6, 0.4, 0.6, 0.4),
                                      (350.0, 25.0, 0.5, 0.5, 0.5, 0.5), (-700.0, 180.0, 1, 1, 1, 1)])
def test_hsvdetector_oracle_equals_numpy(orc, settings):
    px = _px(1 << 20, 2)
    mask = npe.hsvdetector_mask(px[:, 2], px[:, 1], px[:, 0], settings)  # BGRx
    got = orc.hsvdetector(px, 1 << 10, 1 << 10, "BGRx", "RGBA", settings).reshape(-1, 4)
    assert np.array_equal(got[:, 3], mask)
    assert np.array_equal(got[:, 0], px[:, 2]) and np.array_equal(got[:, 2], px[:, 0])
    assert 0 < int((mask == 255).sum()) <= len(mask)


@pytest.mark.parametrize("n,domain", [(33, None), (65, None), (7, ((0.1, 0.0, -0.5), (0.9, 2.0, 0.5)))])
def test_colorlut3d_oracle_equals_numpy(orc, n, domain):
    text = frames.cube_text_3d(n, domain_min=domain[0] if domain else None,
                               domain_max=domain[1] if domain else None)
    lut = orc.Lut(text=text)
    table = lut.data.reshape(-1, 4)
    px = _px(1 << 19, 3)
    want = px.copy()
    want[:, :3] = npe.colorlut_3d(px[:, :3], table, n, lut.scale, lut.offset)
    got = orc.colorlut(lut, px, 1 << 10, 1 << 9).reshape(-1, 4)
    assert np.array_equal(got, want)
    # RGBA64 LE / BE
    raw = frames.random_bytes((1 << 16) * 8, 4)
    for fmt, dt in (("RGBA64_LE", "<u2"), ("RGBA64_BE", ">u2")):
        words = raw.view(dt).reshape(-1, 4)
        out16 = npe.colorlut_3d(words[:, :3].astype(np.uint16), table, n, lut.scale, lut.offset, 65535)
        want16 = words.copy()
        want16[:, :3] = out16
        got16 = orc.colorlut(lut, raw, 1 << 8, 1 << 8, fmt)
        assert np.array_equal(got16, want16.view(np.uint8).reshape(-1)), fmt


def test_colorlut1d_oracle_equals_numpy(orc):
    lut = orc.Lut(text=frames.cube_text_1d(1024, domain_min=(0, 0.25, 0), domain_max=(1, 0.75, 2)))
    planes = lut.data.reshape(3, 1024)
    px = _px(1 << 18, 5)
    want = px.copy()
    want[:, :3] = npe.colorlut_1d(px[:, :3], planes, 1024, lut.scale, lut.offset)
    assert np.array_equal(orc.colorlut(lut, px, 1 << 9, 1 << 9).reshape(-1, 4), want)
