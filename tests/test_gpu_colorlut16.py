"""The RGBA64 fast op (ColorLut64Op: x-differences precomputed at upload, constant strides, /65535
with N-1 folded in, FRND/F2I coordinates) against the oracle: every 16-bit code on every axis, LUT
sizes on both sides of its limits (power-of-two and other N-1, the 65 / 129 stride classes, N > 128
which keeps the direct kernel), LE / BE, LUT values outside [0,1] and non-finite, padded strides."""
import numpy as np
import pytest

import util
import gst_plugins_rs_b200 as g
from gst_plugins_rs_b200 import frames

pytestmark = pytest.mark.gpu


def _all_codes_frame(seed, axis):
    """65536 pixels: channel `axis` runs over every 16-bit code, the others are random."""
    rng = np.random.default_rng(seed)
    px = rng.integers(0, 65536, size=(65536, 4), dtype=np.uint16)
    px[:, axis] = np.arange(65536, dtype=np.uint16)
    return px


@pytest.mark.parametrize("n,fmt", [(n, "RGBA64_LE") for n in (2, 3, 5, 17, 20, 33, 64, 65, 66, 128)] +
                         [(n, "RGBA64_BE") for n in (3, 33, 65)])
def test_fast_op_every_code_on_every_axis(ctx, orc, n, fmt):
    rng = np.random.default_rng(n)
    vals = rng.uniform(0.0, 1.0, size=(n ** 3, 3))
    text = frames.cube_text_3d(n, vals)
    lut = orc.Lut(text=text)
    ctx.set_lut_from_cube(g.parse_cube(text))
    w, h = 256, 256
    for axis in range(3):
        px = _all_codes_frame(100 * n + axis, axis)
        src = px.astype(">u2" if fmt.endswith("BE") else "<u2").view(np.uint8).reshape(-1)
        got = util.gpu_colorlut(ctx, src, w, h, fmt)
        assert ctx.get_option("lut.path_active") == 7
        assert np.array_equal(got, orc.colorlut(lut, src, w, h, fmt)), (n, fmt, axis)


def test_fast_op_non_unit_and_non_finite_lut_values(ctx, orc):
    n = 9
    rng = np.random.default_rng(3)
    vals = rng.uniform(-0.5, 1.5, size=(n ** 3, 3))
    vals[5] = [np.nan, 2.0, -1.0]
    vals[100] = [np.inf, -np.inf, 0.5]
    text = frames.cube_text_3d(n, vals)
    lut = orc.Lut(text=text)
    ctx.set_lut_from_cube(g.parse_cube(text))
    w, h = 640, 96
    for k, fmt in enumerate(("RGBA64_LE", "RGBA64_BE")):
        src = frames.random_bytes(w * h * 8, 50 + k)
        got = util.gpu_colorlut(ctx, src, w, h, fmt)
        assert ctx.get_option("lut.path_active") == 7
        assert np.array_equal(got, orc.colorlut(lut, src, w, h, fmt))


def test_fast_op_limits_and_strides(ctx, orc):
    w, h = 333, 21
    # N = 129 is beyond the op's table classes; a non-identity domain is too: both keep the direct kernel
    for n, dom, want_path in ((129, None, 0), (17, ((0.1, 0.0, 0.0), (0.9, 1.0, 1.0)), 0), (17, None, 7)):
        text = frames.cube_text_3d(n, None, *(dom or (None, None)))
        lut = orc.Lut(text=text)
        ctx.set_lut_from_cube(g.parse_cube(text))
        stride = w * 8 + 24
        src = frames.random_bytes(stride * h, n)
        got = util.gpu_colorlut(ctx, src, w, h, "RGBA64_LE", src_stride=stride, dst_stride=stride + 8)
        assert ctx.get_option("lut.path_active") == want_path
        want = orc.colorlut(lut, src, w, h, "RGBA64_LE", src_stride=stride, dst_stride=stride + 8,
                            dst=np.full(h * (stride + 8), 0xA5, np.uint8))
        assert np.array_equal(got, want), (n, dom)
    ctx.set_option("lut.path", 1)   # forced direct kernel
    src = frames.random_bytes(w * h * 8, 7)
    got = util.gpu_colorlut(ctx, src, w, h, "RGBA64_LE")
    assert ctx.get_option("lut.path_active") == 0
    assert np.array_equal(got, orc.colorlut(lut, src, w, h, "RGBA64_LE"))
