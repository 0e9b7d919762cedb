"""EXTENSION modes without a reference counterpart: tetrahedral and nearest 3D-LUT
interpolation ("lut.interpolation" = 1 / 2).  BASELINE.json names them; gst-plugins-rs implements
trilinear only, so there is no reference parity to claim.  The oracle's definition
(oracle/vf_oracle.c, sample_3d_tetrahedral / sample_3d_nearest) is pinned here against an independent
numpy restatement and against properties of the published algorithm; the CUDA kernels must match
the oracle bit for bit."""
import numpy as np
import pytest

import util
from gst_plugins_rs_b200 import frames

F = np.float32


def _np_coords(lut, codes, maxv):
    """norm_comp (imp.rs:471-479) * (size - 1), in f32."""
    v = codes.astype(F) / F(maxv)
    n = np.clip(v * lut.scale.astype(F)[None, :] + lut.offset.astype(F)[None, :], F(0), F(1)).astype(F)
    return (n * (F(lut.size) - F(1))).astype(F)


def _np_tetrahedral(lut, rgb, maxv):
    """Sort-based formulation (max/mid/min of the fractions pick the path along the cell edges),
    independent of the oracle's six-way branch.  Where fractions tie the two formulations may walk
    different edges, but the corner they disagree on then has weight zero (finite LUT entries)."""
    n = lut.size
    table = lut.data.reshape(n, n, n, 4)[..., :3]          # [z][y][x]
    p = _np_coords(lut, rgb, maxv)
    i0 = np.minimum(np.floor(p).astype(np.int64), n - 1)
    t = (p - i0.astype(F)).astype(F)
    order = np.argsort(-t, axis=1, kind="stable")           # axis of max, mid, min fraction
    ts = np.take_along_axis(t, order, 1)
    w = np.stack([F(1) - ts[:, 0], ts[:, 0] - ts[:, 1], ts[:, 1] - ts[:, 2], ts[:, 2]], 1).astype(F)
    corner = i0.copy()
    acc = None
    for k in range(4):
        c = np.minimum(corner, n - 1)
        val = table[c[:, 2], c[:, 1], c[:, 0]].astype(F)
        term = (w[:, k:k + 1] * val).astype(F)
        acc = term if acc is None else (acc + term).astype(F)
        if k < 3:
            corner = corner.copy()
            corner[np.arange(len(corner)), order[:, k]] += 1
    out = np.floor(np.clip(acc, F(0), F(1)) * F(maxv) + F(0.5))   # round half away, values >= 0
    return out.astype(np.uint16 if maxv > 255 else np.uint8), t


def _np_nearest(lut, rgb, maxv):
    n = lut.size
    table = lut.data.reshape(n, n, n, 4)[..., :3]
    p = _np_coords(lut, rgb, maxv)
    i = np.minimum(np.floor((p + F(0.5)).astype(F)).astype(np.int64), n - 1)
    val = table[i[:, 2], i[:, 1], i[:, 0]].astype(F)
    out = np.floor(np.clip(val, F(0), F(1)) * F(maxv) + F(0.5))
    return out.astype(np.uint16 if maxv > 255 else np.uint8)


def _rand_lut_text(n, seed, domain=None):
    rng = np.random.default_rng(seed)
    vals = rng.random((n ** 3, 3))
    return frames.cube_text_3d(n, vals, *(domain or (None, None)))


@pytest.mark.parametrize("n", [2, 5, 17, 33])
def test_oracle_tetrahedral_and_nearest_vs_numpy(orc, n):
    lut = orc.Lut(text=_rand_lut_text(n, n, ((0.0, 0.1, 0.0), (1.0, 0.9, 0.8)) if n == 5 else None))
    w, h = 257, 64
    src = frames.frame_rand(w, h, 4, n).reshape(-1, 4)
    got_t = orc.colorlut(lut, src.reshape(-1), w, h, interpolation="tetrahedral").reshape(-1, 4)
    got_n = orc.colorlut(lut, src.reshape(-1), w, h, interpolation="nearest").reshape(-1, 4)
    want_t, t = _np_tetrahedral(lut, src[:, :3], 255)
    distinct = (t[:, 0] != t[:, 1]) & (t[:, 1] != t[:, 2]) & (t[:, 0] != t[:, 2])
    assert distinct.mean() > 0.5 and not distinct.all()   # ties occur (e.g. equal channel codes) ...
    assert np.array_equal(got_t[:, :3], want_t)           # ... and cost nothing: their weight is zero
    assert np.array_equal(got_n[:, :3], _np_nearest(lut, src[:, :3], 255))
    assert np.array_equal(got_t[:, 3], src[:, 3]) and np.array_equal(got_n[:, 3], src[:, 3])
    # 16-bit, little endian
    src16 = np.frombuffer(frames.frame_rand(w, h, 8, n + 1).tobytes(), "<u2").reshape(-1, 4)
    got16 = np.frombuffer(orc.colorlut(lut, src16.view(np.uint8).reshape(-1), w, h, "RGBA64_LE",
                                       interpolation="tetrahedral").tobytes(), "<u2").reshape(-1, 4)
    want16, t16 = _np_tetrahedral(lut, src16[:, :3], 65535)
    d16 = (t16[:, 0] != t16[:, 1]) & (t16[:, 1] != t16[:, 2]) & (t16[:, 0] != t16[:, 2])
    assert d16.mean() > 0.5
    assert np.array_equal(got16[:, :3], want16) and np.array_equal(got16[:, 3], src16[:, 3])


def test_oracle_extension_properties(orc):
    # identity LUT: every mode is the identity on 8-bit codes (nearest needs one node per code)
    src = frames.all_rgb_frame().reshape(-1)   # all 2^24 RGB triples
    w, h = 4096, 4096
    for n in (2, 17, 33):
        lut = orc.Lut(text=frames.cube_text_3d(n, frames.identity_lut_values(n)))
        assert np.array_equal(orc.colorlut(lut, src, w, h, interpolation="tetrahedral"), src)
    # nearest on an identity LUT whose nodes are the multiples of 5 (N = 52): nodes map to themselves,
    # every other code snaps to the closest node (ties cannot occur: 2.5 is not a code distance)
    lut = orc.Lut(text=frames.cube_text_3d(52, frames.identity_lut_values(52)))
    small = frames.frame_rand(256, 64, 4, 3)
    got = orc.colorlut(lut, small, 256, 64, interpolation="nearest").reshape(-1, 4)
    want = ((small.reshape(-1, 4)[:, :3].astype(np.int32) + 2) // 5 * 5).astype(np.uint8)
    assert np.array_equal(got[:, :3], want)
    # a constant LUT is constant under every mode
    lut = orc.Lut(text=frames.cube_text_3d(7, np.tile([[0.25, 0.5, 0.75]], (343, 1))))
    for mode in ("trilinear", "tetrahedral", "nearest"):
        out = orc.colorlut(lut, small, 256, 64, interpolation=mode).reshape(-1, 4)
        assert (out[:, :3] == np.array([64, 128, 191], np.uint8)).all(), mode
    node = small
    # grey axis r=g=b of a size-2 LUT: tetrahedral walks the main diagonal c000 → c111 only
    vals = np.zeros((8, 3))
    vals[7] = (1.0, 0.5, 0.25)            # c111; every other corner black
    vals[1] = vals[2] = vals[4] = (1, 1, 1)  # would show up if an edge corner were weighted
    lut = orc.Lut(text=frames.cube_text_3d(2, vals))
    grey = np.repeat(np.arange(256, dtype=np.uint8), 4).reshape(-1, 4).copy()
    out = orc.colorlut(lut, grey.reshape(-1), 256, 1, interpolation="tetrahedral").reshape(-1, 4)
    t = np.arange(256, dtype=F) / F(255)
    for c, top in enumerate((1.0, 0.5, 0.25)):
        want = np.floor(((F(1) - t) * F(0) + t * F(top)).astype(F) * F(255) + F(0.5)).astype(np.uint8)
        assert np.array_equal(out[:, c], want)
    # 1D LUTs ignore the mode
    lut1 = orc.Lut(text=frames.cube_text_1d(16))
    assert np.array_equal(orc.colorlut(lut1, node, 64, 64, interpolation="tetrahedral"),
                          orc.colorlut(lut1, node, 64, 64))


# --------------------------------------------------------------------------------------
def _load(ctx, orc, text):
    import gst_plugins_rs_b200 as g
    ctx.set_lut_from_cube(g.parse_cube(text))
    return orc.Lut(text=text)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["tetrahedral", "nearest"])
@pytest.mark.parametrize("n", [2, 3, 17, 33, 65])
def test_gpu_extension_modes_match_oracle(ctx, orc, mode, n):
    lut = _load(ctx, orc, frames.cube_text_3d(n))
    ctx.set_option("lut.interpolation", {"tetrahedral": 1, "nearest": 2}[mode])
    assert ctx.get_option("lut.interpolation") in (1, 2)
    w, h = 1280, 128
    for lut_path in (0, 1, 4):  # auto (= baked for these modes), direct kernel, baked
        ctx.set_option("lut.path", lut_path)
        for name, src in (("bars", frames.frame_bars(w, h)), ("grad", frames.frame_grad(w, h)),
                          ("rand", frames.frame_rand(w, h, 4, 11))):
            got = util.gpu_colorlut(ctx, src, w, h)
            want = orc.colorlut(lut, src, w, h, interpolation=mode)
            mx, exact = util.diff_report(got, want)
            assert mx == 0 and exact == 1.0, f"{mode} {name} n={n} path={lut_path}: {mx} {exact:.6f}"
    # back to the reference mode: the baked table is rebuilt, results are trilinear again
    ctx.set_option("lut.interpolation", 0)
    src = frames.frame_rand(w, h, 4, 12)
    assert np.array_equal(util.gpu_colorlut(ctx, src, w, h), orc.colorlut(lut, src, w, h))


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["tetrahedral", "nearest"])
def test_gpu_extension_modes_formats_domain_and_odd_geometry(ctx, orc, mode):
    """RGBA64 LE/BE, non-identity domain (incl. clamped coordinates), random LUT values, odd
    width with padded strides, host frames."""
    text = _rand_lut_text(9, 4, ((0.1, 0.0, 0.2), (0.9, 1.0, 0.7)))
    lut = _load(ctx, orc, text)
    ctx.set_option("lut.interpolation", {"tetrahedral": 1, "nearest": 2}[mode])
    w, h = 333, 41
    for fmt, bpp in (("RGBA", 4), ("RGBA64_LE", 8), ("RGBA64_BE", 8)):
        stride = w * bpp + 24
        src = np.zeros((h, stride), np.uint8)
        src[:, :w * bpp] = frames.frame_rand(w, h, bpp, 21).reshape(h, w * bpp)
        for memory in ("device", "host"):
            got = util.gpu_colorlut(ctx, src.reshape(-1), w, h, fmt, stride, stride, memory=memory)
            want = orc.colorlut(lut, src.reshape(-1), w, h, fmt, stride, stride,
                                dst=np.full(h * stride, 0xA5, np.uint8), interpolation=mode)
            assert np.array_equal(got, want), f"{mode} {fmt} {memory}"


@pytest.mark.gpu
def test_gpu_chain_with_extension_mode_equals_two_passes(ctx, orc):
    import gst_plugins_rs_b200 as g
    from gst_plugins_rs_b200.api import frame_of
    import torch
    lut = _load(ctx, orc, frames.cube_text_3d(17))
    ctx.set_option("lut.interpolation", 1)
    w, h = 640, 64
    src = frames.frame_rand(w, h, 4, 5)
    tin = torch.from_numpy(src.copy()).cuda()
    tout = torch.zeros_like(tin)
    ctx.chain_lut_hsv_batch([frame_of(tin, w, h, "RGBA")], [frame_of(tout, w, h, "RGBA")],
                            g.HsvFilterParams(*util.CFG2))
    ctx.synchronize()
    want = orc.hsvfilter(orc.colorlut(lut, src, w, h, interpolation="tetrahedral"), w, h, "RGBA", util.CFG2)
    assert np.array_equal(tout.cpu().numpy().reshape(-1), np.asarray(want).reshape(-1))
