"""GPU parity: hsvfilter / hsvdetector through the C ABI vs the CPU oracle.

Bar (integer/byte outputs): bit-exact.  hsvdetector: 0 mask mismatches.
The 2^24-triple sweeps make the RGB→HSV→RGB shortcuts of vf_math.cuh a proof by
enumeration for each parameter set.
"""
import numpy as np
import pytest

import util
from gst_plugins_rs_b200 import frames

pytestmark = pytest.mark.gpu

FILTER_SETTINGS = [
    util.IDENTITY,
    util.CFG2,
    (-123.25, 0.7, -0.1, 1.3, 0.1),       # negative shift, clamps on both sides
    (360.0, 1.0, 0.0, 1.0, 0.0),          # edge of the non-negative-shift variant
    (-0.0, 0.8, 0.1, 1.1, -0.05),         # zero-shift variant (hue untouched)
    (-17.0, 1.0, 0.0, 1.0, 0.0),          # negative-shift variant
    (-360.0, 2.5, 0.5, 0.25, 0.5),
    (1234.5, 1.0, 0.25, 1.0, -0.25),      # generic fmod variant
    (-100000.0, 1.1, 0.0, 0.9, 0.0),
]


@pytest.mark.parametrize("settings", FILTER_SETTINGS)
@pytest.mark.parametrize("math", [0, 1])
def test_hsvfilter_exhaustive_rgba(ctx, orc, settings, math):
    """All 2^24 RGB triples, RGBA, device-resident (hsvfilter/imp.rs:76-120)."""
    ctx.set_option("hsv.math", math)
    src = frames.all_rgb_frame(0, 1, 2, 3, other_value=77)
    got = util.gpu_hsvfilter(ctx, src, 4096, 4096, "RGBA", settings)
    want = orc.hsvfilter(src, 4096, 4096, "RGBA", settings)
    mx, exact = util.diff_report(got, want)
    assert mx == 0 and exact == 1.0, f"max diff {mx}, exact fraction {exact:.6f}"


@pytest.mark.parametrize("math", [0, 1])
def test_from_rgb_floats_exhaustive(ctx, orc, math):
    """RGB→HSV as FLOATS, all 2^24 inputs, bit for bit against the oracle (hsvutils.rs:44-84):
    a proof by enumeration that the fast device function (reciprocal-based divisions, predicated
    arm selection, dropped no-op clamps / fmod) equals the reference for its whole domain."""
    import torch
    from gst_plugins_rs_b200.api import debug_hsv_from_rgb
    ctx.set_option("hsv.math", math)
    src = frames.all_rgb_frame().reshape(-1, 4)
    t = torch.from_numpy(src.reshape(-1)).cuda()
    hsv = torch.empty(src.shape[0] * 3, dtype=torch.float32, device="cuda")
    debug_hsv_from_rgb(ctx, t, hsv)
    ctx.synchronize()
    got = hsv.cpu().numpy().view(np.uint32).reshape(-1, 3)
    want = orc.from_rgba_batch(src).view(np.uint32)
    bad = np.nonzero((got != want).any(1))[0]
    assert bad.size == 0, f"{bad.size} triples differ, first {src[bad[0]]}: {got[bad[0]]} vs {want[bad[0]]}"


def test_hsvfilter_identity_regression_fact(ctx):
    """SURVEY.md §8c probe (i): identity settings change 11,093,274 of 2^24 triples, each by 1."""
    src = frames.all_rgb_frame()
    got = util.gpu_hsvfilter(ctx, src, 4096, 4096, "RGBA", util.IDENTITY).reshape(-1, 4)
    s = src.reshape(-1, 4)
    d = np.abs(got[:, :3].astype(np.int16) - s[:, :3].astype(np.int16))
    assert int((d != 0).any(1).sum()) == 11093274
    assert int(d.max()) == 1
    assert (got[:, 3] == 77).all()


@pytest.mark.parametrize("settings", [
    (float("nan"), 1.0, 0.0, 1.0, 0.0),
    (float("inf"), 1.0, 0.0, 1.0, 0.0),
    (10.0, float("nan"), 0.0, 1.0, 0.0),
    (10.0, 1.0, float("inf"), float("-inf"), 0.5),
    (3.0e38, 1.0, 0.0, 1.0, 0.0),
    (1e-30, 1e30, -1e30, 1.0, float("nan")),
])
def test_hsvfilter_nonfinite_settings(ctx, orc, settings):
    """Property range is ±3.4e38 and GObject lets NaN/inf through: match the Rust semantics
    (Clamp trait NaN→0, NaN hue → (m,m,m); hsvfilter/imp.rs:102-115, hsvutils.rs:138-154)."""
    src = frames.frame_rand(1024, 256, 4, frame_index=3)
    for math in (0, 1):
        ctx.set_option("hsv.math", math)
        got = util.gpu_hsvfilter(ctx, src, 1024, 256, "RGBA", settings)
        want = orc.hsvfilter(src, 1024, 256, "RGBA", settings)
        assert util.diff_report(got, want) == (0, 1.0)


FORMATS10 = ["RGBx", "xRGB", "BGRx", "xBGR", "RGBA", "ARGB", "BGRA", "ABGR", "RGB", "BGR"]


@pytest.mark.parametrize("fmt", FORMATS10)
@pytest.mark.parametrize("memory", ["device", "host"])
def test_hsvfilter_formats_strides(ctx, orc, fmt, memory):
    """All 10 caps formats (hsvfilter/imp.rs:278-289, arms 327-371); odd width, padded stride,
    misaligned base; padding bytes must stay untouched (imp.rs:94-97)."""
    bpp = 3 if fmt in ("RGB", "BGR") else 4
    for (w, h, pad, off) in [(253, 37, 0, 0), (640, 16, 64, 0), (101, 9, 5, 3), (1, 1, 0, 0),
                             (4, 3, 16, 16)]:
        stride = w * bpp + pad
        if bpp == 3:
            stride = (stride + 3) & ~3  # GStreamer rounds RGB strides up to 4
        total = off + stride * h
        buf = frames.random_bytes(total, frame_index=w + h)
        want = buf.copy()
        want[off:] = orc.hsvfilter(buf[off:], w, h, fmt, util.CFG2, stride=stride)
        import torch
        from gst_plugins_rs_b200.api import frame_of
        import gst_plugins_rs_b200 as g
        if memory == "device":
            t = torch.from_numpy(buf.copy()).cuda()
        else:
            t = buf.copy()
        ctx.hsvfilter(frame_of(t, w, h, fmt, stride, offset=off), g.HsvFilterParams(*util.CFG2))
        ctx.synchronize()
        got = t.cpu().numpy() if memory == "device" else t
        assert np.array_equal(got, want), f"{fmt} {w}x{h} pad {pad} off {off} ({memory})"


DET_SETTINGS = [
    util.DET_DEFAULT,
    util.DET_CFG4,
    (350.0, 25.0, 0.5, 0.5, 0.5, 0.5),     # hue window wrapping through 0 (negative offset)
    (180.0, 45.0, 0.5, 0.5, 0.5, 0.5),     # offset exactly 0
    (540.0, 90.0, 0.5, 0.5, 0.5, 0.5),     # offset -360: the edge of the negative variant
    (-180.0, 90.0, 0.5, 0.5, 0.5, 0.5),    # offset +360: the edge of the non-negative variant
    (-700.0, 180.0, 1.0, 1.0, 1.0, 1.0),   # generic fmod variant, everything matches on s/v
    (45.0, 0.0, 0.25, 0.0, 0.75, 0.0),     # zero-width windows: exact-equality thresholds
]


@pytest.mark.parametrize("settings", DET_SETTINGS)
@pytest.mark.parametrize("math", [0, 1])
def test_hsvdetector_exhaustive_bgrx(ctx, orc, settings, math):
    """All 2^24 triples as BGRx → RGBA (cfg4's format pair): zero mask mismatches and exact
    colour copies (hsvdetector/imp.rs:100-160)."""
    ctx.set_option("hsv.math", math)
    src = frames.all_rgb_frame(2, 1, 0, 3, other_value=9)  # B,G,R,x
    got = util.gpu_hsvdetector(ctx, src, 4096, 4096, "BGRx", "RGBA", settings).reshape(-1, 4)
    want = orc.hsvdetector(src, 4096, 4096, "BGRx", "RGBA", settings).reshape(-1, 4)
    mism = int((got[:, 3] != want[:, 3]).sum())
    assert mism == 0, f"{mism} mask mismatches"
    assert np.array_equal(got[:, :3], want[:, :3])
    n_match = int((want[:, 3] == 255).sum())
    assert set(np.unique(want[:, 3])) <= {0, 255}
    print("matching pixels:", n_match)


IN6 = ["RGBx", "xRGB", "BGRx", "xBGR", "RGB", "BGR"]
OUT4 = ["RGBA", "ARGB", "BGRA", "ABGR"]


@pytest.mark.parametrize("in_fmt", IN6)
@pytest.mark.parametrize("out_fmt", OUT4)
def test_hsvdetector_all_24_pairs(ctx, orc, in_fmt, out_fmt):
    """6 sink × 4 src formats (hsvdetector/imp.rs:78-96, closures 428-704)."""
    bpp = 3 if in_fmt in ("RGB", "BGR") else 4
    for memory in ("device", "host"):
        for (w, h, pad) in [(317, 21, 0), (64, 8, 32)]:
            in_stride = (w * bpp + pad + 3) & ~3
            out_stride = w * 4 + pad
            src = frames.random_bytes(in_stride * h, frame_index=7)
            got = util.gpu_hsvdetector(ctx, src, w, h, in_fmt, out_fmt, util.DET_CFG4, in_stride,
                                       out_stride, memory=memory)
            want = orc.hsvdetector(src, w, h, in_fmt, out_fmt, util.DET_CFG4, in_stride,
                                   out_stride, dst=np.full(h * out_stride, 0xA5, np.uint8))
            assert np.array_equal(got, want), f"{in_fmt}->{out_fmt} {w}x{h} ({memory})"


def test_hsvfilter_batch_and_host_pinned(ctx, orc):
    """process_batch over device frames and a pinned host frame through the stream pipeline."""
    import torch
    import gst_plugins_rs_b200 as g
    from gst_plugins_rs_b200.api import frame_of
    w, h = 1920, 1080
    srcs = [frames.frame_of_class(c, w, h, i) for i, c in enumerate(["bars", "grad", "rand"] * 12)]
    ts = [torch.from_numpy(s.reshape(-1).copy()).cuda() for s in srcs]
    ctx.hsvfilter_batch([frame_of(t, w, h, "RGBA") for t in ts], g.HsvFilterParams(*util.CFG2))
    ctx.synchronize()
    wants = {}
    for i, (s, t) in enumerate(zip(srcs, ts)):
        key = i % 3 if i % 3 < 2 else i
        if key not in wants:
            wants[key] = orc.hsvfilter(s, w, h, "RGBA", util.CFG2)
        assert np.array_equal(t.cpu().numpy(), wants[key]), f"batch frame {i}"
    # pinned host frame (direct 2D async copies, chunked)
    p = torch.from_numpy(srcs[2].reshape(-1).copy()).pin_memory()
    ctx.hsvfilter(frame_of(p, w, h, "RGBA"), g.HsvFilterParams(*util.CFG2))
    assert np.array_equal(p.numpy(), wants[2])
    st = ctx.stats()
    assert st["h2d_bytes"] == w * h * 4 and st["d2h_bytes"] == w * h * 4
    assert st["kernel_launches"] >= 3
