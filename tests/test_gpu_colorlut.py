"""GPU parity: colorlut through the C ABI vs the CPU oracle (colorlut/imp.rs:203-543).

Both LUT paths (direct 8-corner and R-resampled table) and both math variants use the
reference's f32 operation order, so the bar is bit-exact on every 8/16-bit output.
"""
import os

import numpy as np
import pytest

import util
from gst_plugins_rs_b200 import frames
from gst_plugins_rs_b200.api import B200VFError, ERR_NO_LUT, ERR_SETTINGS, ERR_IO, ERR_PARSE

pytestmark = pytest.mark.gpu


def _load(ctx, orc, text):
    lut = orc.Lut(text=text)
    import gst_plugins_rs_b200 as g
    ctx.set_lut_from_cube(g.parse_cube(text))
    return lut


def _test_frames(w, h):
    yield "bars", frames.frame_bars(w, h)
    yield "grad", frames.frame_grad(w, h)
    yield "rand", frames.frame_rand(w, h, 4, 11)


@pytest.mark.parametrize("n", [2, 3, 17, 33, 65])
@pytest.mark.parametrize("lut_path", [1, 2, 3, 4])
@pytest.mark.parametrize("math", [0, 1])
def test_colorlut_3d_rgba(ctx, orc, n, lut_path, math):
    """Synthetic §8(d) LUT, trilinear, RGBA, three content classes."""
    lut = _load(ctx, orc, frames.cube_text_3d(n))
    ctx.set_option("lut.path", lut_path)
    ctx.set_option("hsv.math", math)
    w, h = 1280, 256
    for name, src in _test_frames(w, h):
        got = util.gpu_colorlut(ctx, src, w, h)
        want = orc.colorlut(lut, src, w, h)
        mx, exact = util.diff_report(got, want)
        assert mx == 0 and exact == 1.0, f"{name} n={n}: max diff {mx}, exact {exact:.6f}"


@pytest.mark.parametrize("n", [2, 17, 33, 64, 65, 256])
def test_colorlut_identity_lut_is_identity(ctx, orc, n):
    """Invariant: an identity .cube reproduces every 8-bit code exactly."""
    if n == 256:
        vals = frames.identity_lut_values(n)
    else:
        vals = frames.identity_lut_values(n)
    _load(ctx, orc, frames.cube_text_3d(n, vals))
    w, h = 2048, 128
    src = frames.frame_rand(w, h, 4, 5)
    src.reshape(-1, 4)[:256, :3] = np.arange(256, dtype=np.uint8)[:, None]  # all greys
    for lut_path in (1, 2, 3, 4):
        ctx.set_option("lut.path", lut_path)
        got = util.gpu_colorlut(ctx, src, w, h)
        assert np.array_equal(got, src.reshape(-1)), f"identity LUT n={n} path={lut_path}"


def test_colorlut_all_2_24_inputs_33(ctx, orc):
    """Every RGB triple through the 33^3 synthetic LUT (cfg1's LUT), both LUT paths."""
    lut = _load(ctx, orc, frames.cube_text_3d(33))
    src = frames.all_rgb_frame()
    want = orc.colorlut(lut, src, 4096, 4096)
    for lut_path in (1, 2, 3, 4):
        ctx.set_option("lut.path", lut_path)
        got = util.gpu_colorlut(ctx, src, 4096, 4096)
        mx, exact = util.diff_report(got, want)
        assert mx == 0 and exact == 1.0, f"path {lut_path}: max diff {mx}, exact {exact:.7f}"


def test_colorlut_all_2_24_inputs_65_default_path(ctx, orc):
    """BASELINE configs[2]'s LUT size (65^3, trilinear) over every RGB triple, default path."""
    lut = _load(ctx, orc, frames.cube_text_3d(65))
    ctx.set_option("lut.path", 0)
    src = frames.all_rgb_frame()
    got = util.gpu_colorlut(ctx, src, 4096, 4096)
    mx, exact = util.diff_report(got, orc.colorlut(lut, src, 4096, 4096))
    assert mx == 0 and exact == 1.0, f"max diff {mx}, exact {exact:.7f}"


def test_colorlut_max_size_256(ctx, orc):
    """LUT_3D_SIZE 256 (parser.rs:16): every t is 0, a pure lookup (SURVEY.md §8c probe v); the
    RG table is not built at this size, so auto falls back to the R-resampled path."""
    n = 256
    rng = np.random.default_rng(7)
    vals = rng.uniform(0, 1, size=(n ** 3, 3)).astype(np.float32)
    import gst_plugins_rs_b200 as g
    data = np.concatenate([vals, np.ones((n ** 3, 1), np.float32)], 1).reshape(-1)
    ctx.set_lut(3, n, data)
    w, h = 2048, 64
    src = frames.frame_rand(w, h, 4, 6)
    px = src.reshape(-1, 4)
    idx = px[:, 0].astype(np.int64) + px[:, 1].astype(np.int64) * n + px[:, 2].astype(np.int64) * n * n
    want = px.copy()
    want[:, :3] = np.floor(vals[idx] * np.float32(255.0) + np.float32(0.5)).astype(np.uint8)
    for lut_path in (0, 1, 2):
        ctx.set_option("lut.path", lut_path)
        got = util.gpu_colorlut(ctx, src, w, h)
        assert np.array_equal(got, want.reshape(-1)), lut_path


@pytest.mark.parametrize("domain", [((0.1, 0.0, -0.5), (0.9, 2.0, 0.5)),
                                    ((-1.0, -1.0, -1.0), (3.0, 1.5, 1.0))])
def test_colorlut_domain_scaling(ctx, orc, domain):
    """DOMAIN_MIN/MAX → scale/offset (parser.rs:264-274) → clamp (imp.rs:471-474)."""
    text = frames.cube_text_3d(17, domain_min=domain[0], domain_max=domain[1])
    lut = _load(ctx, orc, text)
    w, h = 1024, 64
    src = frames.frame_rand(w, h, 4, 2)
    for lut_path in (1, 2, 3, 4):
        for math in (0, 1):
            ctx.set_option("lut.path", lut_path)
            ctx.set_option("hsv.math", math)
            got = util.gpu_colorlut(ctx, src, w, h)
            want = orc.colorlut(lut, src, w, h)
            assert util.diff_report(got, want) == (0, 1.0)


def test_colorlut_out_of_range_and_nonfinite_entries(ctx, orc):
    """LUT entries outside [0,1] are legal (only the output is clamped, imp.rs:538); inf/nan
    parse too (Rust f32::from_str) and must end as the same bytes."""
    rng = np.random.default_rng(1)
    vals = rng.uniform(-0.5, 1.5, size=(5 ** 3, 3))
    lines = ["LUT_3D_SIZE 5"] + ["%.6f %.6f %.6f" % tuple(v) for v in vals]
    lines[10] = "inf 0.5 -inf"
    lines[40] = "nan 0.25 1e-40"
    lut = _load(ctx, orc, "\n".join(lines) + "\n")
    w, h = 512, 64
    src = frames.frame_rand(w, h, 4, 8)
    want = orc.colorlut(lut, src, w, h)
    for lut_path in (1, 2, 3, 4):
        for math in (0, 1):
            ctx.set_option("lut.path", lut_path)
            ctx.set_option("hsv.math", math)
            got = util.gpu_colorlut(ctx, src, w, h)
            assert util.diff_report(got, want) == (0, 1.0), (lut_path, math)


@pytest.mark.parametrize("n", [2, 256, 1024, 65536])
def test_colorlut_1d_rgba(ctx, orc, n):
    """1D LUT, linear (imp.rs:237-265, 399-413, 482-490)."""
    lut = _load(ctx, orc, frames.cube_text_1d(n))
    w, h = 1024, 64
    src = frames.frame_rand(w, h, 4, 4)
    for math in (0, 1):
        ctx.set_option("hsv.math", math)
        got = util.gpu_colorlut(ctx, src, w, h)
        want = orc.colorlut(lut, src, w, h)
        assert util.diff_report(got, want) == (0, 1.0)


@pytest.mark.parametrize("fmt", ["RGBA64_LE", "RGBA64_BE"])
@pytest.mark.parametrize("kind", ["3d", "1d", "3d-domain"])
def test_colorlut_rgba64(ctx, orc, fmt, kind):
    """RGBA64 LE/BE incl. raw alpha copy (imp.rs:307-397, 415-429, 451-469)."""
    if kind == "3d":
        text = frames.cube_text_3d(33)
    elif kind == "1d":
        text = frames.cube_text_1d(4096)
    else:
        text = frames.cube_text_3d(9, domain_min=(0.0, 0.1, 0.0), domain_max=(1.0, 0.8, 2.0))
    lut = _load(ctx, orc, text)
    for (w, h, pad) in [(640, 32, 0), (333, 7, 16), (5, 3, 2)]:
        stride = w * 8 + pad
        src = frames.random_bytes(stride * h, frame_index=w)
        for memory in ("device", "host"):
            for math in (0, 1):
                ctx.set_option("hsv.math", math)
                got = util.gpu_colorlut(ctx, src, w, h, fmt, stride, stride, memory=memory)
                want = orc.colorlut(lut, src, w, h, fmt, stride, stride,
                                    dst=np.full(h * stride, 0xA5, np.uint8))
                assert np.array_equal(got, want), f"{fmt} {kind} {w}x{h} {memory} math={math}"


def test_colorlut_strides_alignment_padding(ctx, orc):
    """Odd widths, padded and misaligned rows, different in/out strides; padding untouched."""
    lut = _load(ctx, orc, frames.cube_text_3d(33))
    import torch
    from gst_plugins_rs_b200.api import frame_of
    for (w, h, spad, dpad, off) in [(1919, 13, 0, 0, 0), (1000, 9, 64, 128, 0), (77, 5, 3, 9, 1),
                                    (1, 1, 0, 0, 0), (3, 2, 4, 4, 4)]:
        ss, ds = w * 4 + spad, w * 4 + dpad
        src = frames.random_bytes(off + ss * h, frame_index=w)
        dst0 = np.full(off + ds * h, 0x5A, np.uint8)
        want = dst0.copy()
        orc.colorlut(lut, src[off:], w, h, "RGBA", ss, ds, dst=want[off:])
        for memory in ("device", "host"):
            s = torch.from_numpy(src.copy()).cuda() if memory == "device" else src.copy()
            d = torch.from_numpy(dst0.copy()).cuda() if memory == "device" else dst0.copy()
            ctx.colorlut(frame_of(s, w, h, "RGBA", ss, offset=off),
                         frame_of(d, w, h, "RGBA", ds, offset=off))
            ctx.synchronize()
            got = d.cpu().numpy() if memory == "device" else d
            assert np.array_equal(got, want), f"{w}x{h} spad={spad} dpad={dpad} off={off} {memory}"


def test_colorlut_errors_and_lifecycle(ctx, tmp_path):
    """start/stop error behaviour (colorlut/imp.rs:168-199, 210-213)."""
    import torch
    from gst_plugins_rs_b200.api import frame_of
    t = torch.zeros(64 * 4, dtype=torch.uint8, device="cuda")
    f = frame_of(t, 64, 1, "RGBA")
    with pytest.raises(B200VFError) as e:
        ctx.colorlut(f, f)
    assert e.value.status == ERR_NO_LUT and "No LUT configured" in e.value.message
    with pytest.raises(B200VFError) as e:
        ctx.set_lut_file(None)
    assert e.value.status == ERR_SETTINGS
    with pytest.raises(B200VFError) as e:
        ctx.set_lut_file(tmp_path / "missing.cube")
    assert e.value.status == ERR_IO and "Failed to parse LUT file" in e.value.message
    bad = tmp_path / "bad.cube"
    bad.write_text("LUT_3D_SIZE 2\n0 0 0\n")
    with pytest.raises(B200VFError) as e:
        ctx.set_lut_file(bad)
    assert e.value.status == ERR_PARSE
    good = tmp_path / "good.cube"
    good.write_text(frames.cube_text_3d(4))
    ctx.set_lut_file(good)
    ctx.colorlut(f, f)
    ctx.clear_lut()  # stop
    with pytest.raises(B200VFError):
        ctx.colorlut(f, f)


def test_chain_equals_two_elements(ctx, orc):
    """Fused colorlut ! hsvfilter == colorlut then hsvfilter."""
    import torch
    import gst_plugins_rs_b200 as g
    from gst_plugins_rs_b200.api import frame_of
    lut = _load(ctx, orc, frames.cube_text_3d(33))
    w, h = 1920, 270
    for name, src in _test_frames(w, h):
        want = orc.hsvfilter(orc.colorlut(lut, src, w, h), w, h, "RGBA", util.CFG2)
        for lut_path in (1, 2, 3, 4):
            ctx.set_option("lut.path", lut_path)
            s = torch.from_numpy(src.reshape(-1).copy()).cuda()
            d = torch.zeros_like(s)
            ctx.chain_lut_hsv_batch([frame_of(s, w, h, "RGBA")], [frame_of(d, w, h, "RGBA")],
                                    g.HsvFilterParams(*util.CFG2))
            ctx.synchronize()
            assert np.array_equal(d.cpu().numpy(), want), f"chain {name} path {lut_path}"
    # default path (baked LUT stage inside the fused kernel) for every hue-shift kernel variant
    ctx.set_option("lut.path", 0)
    src = frames.frame_rand(w, h, 4, 3)
    mid = orc.colorlut(lut, src, w, h)
    for hue in (0.0, -75.0, 400.0, -1000.5):
        params = (hue, 0.8, 0.1, 1.1, -0.05)
        s = torch.from_numpy(src.reshape(-1).copy()).cuda()
        d = torch.zeros_like(s)
        ctx.chain_lut_hsv_batch([frame_of(s, w, h, "RGBA")], [frame_of(d, w, h, "RGBA")],
                                g.HsvFilterParams(*params))
        ctx.synchronize()
        want = orc.hsvfilter(mid.copy(), w, h, "RGBA", params)
        assert np.array_equal(d.cpu().numpy(), want), f"chain hue {hue}"
    # rows that are not 16-byte aligned (padded stride 4*w + 4): served by two launches, same bytes
    w2, h2 = 333, 17
    stride = w2 * 4 + 4
    raw = frames.random_bytes(stride * h2, 77)
    want2 = np.zeros(stride * h2, np.uint8)
    orc.colorlut(lut, raw, w2, h2, "RGBA", stride, stride, dst=want2)
    want2 = orc.hsvfilter(want2, w2, h2, "RGBA", util.CFG2, stride=stride)
    s2 = torch.from_numpy(raw.copy()).cuda()
    d2 = torch.zeros_like(s2)
    ctx.set_option("lut.path", 0)
    ctx.chain_lut_hsv_batch([frame_of(s2, w2, h2, "RGBA", stride)], [frame_of(d2, w2, h2, "RGBA", stride)],
                            g.HsvFilterParams(*util.CFG2))
    ctx.synchronize()
    assert np.array_equal(d2.cpu().numpy(), want2)
