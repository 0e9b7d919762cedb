"""BASELINE.json's full sizes, checked through properties that do not need the (slow) CPU oracle:
batches of 64 4K frames (cfg2 / cfg3 / cfg4) and 8K frames (cfg5) stay on the device and are
compared with torch on the GPU.

* identity LUT: every byte comes back (SURVEY.md §8c invariant), in every batch slot;
* two independent implementations agree on every byte at full size: the function table against
  the per-pixel kernel (hsvfilter, hsvdetector), the baked LUT table against the interpolating
  kernels (colorlut), the fused chain against the two element passes;
* hsvdetector keeps the colour bytes and writes only 0 / 255 alpha; identity hsvfilter changes a
  byte by at most one code and never the alpha byte (SURVEY.md probe fact);
* a crop of the full-size result equals the oracle's (ties the properties to the reference port).
"""
import numpy as np
import pytest
import torch

import util
import gst_plugins_rs_b200 as g
from gst_plugins_rs_b200 import frames
from gst_plugins_rs_b200.api import frame_array, frame_of

pytestmark = pytest.mark.gpu

W4K, H4K = 3840, 2160
W8K, H8K = 7680, 4320


def _batch(n, w, h, classes=("noise", "grad", "bars", "rand")):
    """n device frames cycling through the content classes (8 distinct frames, cloned)."""
    base = [torch.from_numpy(np.ascontiguousarray(frames.frame_of_class(classes[i % len(classes)], w, h, i))
                             .reshape(-1)).cuda() for i in range(min(n, 8))]
    return [base[i % len(base)].clone() for i in range(n)]


def _ctx():
    """A context working on torch's current stream: the torch ops that prepare and compare the
    frames (clone, zero_, equal) and the library's kernels are then ordered with each other."""
    ctx = g.Context(0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    return ctx


def _frames(ts, w, h, fmt):
    return frame_array([frame_of(t, w, h, fmt) for t in ts])


def test_identity_lut_reproduces_64_4k_frames():
    with _ctx() as ctx:
        ctx.set_lut_from_cube(g.parse_cube(frames.cube_text_3d(33, frames.identity_lut_values(33))))
        src = _batch(64, W4K, H4K)
        dst = [torch.zeros_like(t) for t in src]
        for path in (0, 1):   # auto (baked table) and the direct kernel
            ctx.set_option("lut.path", path)
            for d in dst:
                d.zero_()
            ctx.colorlut_batch(_frames(src, W4K, H4K, "RGBA"), _frames(dst, W4K, H4K, "RGBA"))
            ctx.synchronize()
            assert all(torch.equal(a, b) for a, b in zip(src, dst)), f"lut.path={path}"


def test_colorlut65_table_and_interpolating_kernels_agree_on_64_4k_frames(orc):
    text = frames.cube_text_3d(65)
    with _ctx() as ctx:
        ctx.set_lut_from_cube(g.parse_cube(text))
        src = _batch(64, W4K, H4K)
        outs = {}
        for path in (4, 1, 3):   # baked table, direct 8-corner, RG-resampled
            ctx.set_option("lut.path", path)
            dst = [torch.zeros_like(t) for t in src]
            ctx.colorlut_batch(_frames(src, W4K, H4K, "RGBA"), _frames(dst, W4K, H4K, "RGBA"))
            ctx.synchronize()
            outs[path] = dst
        for path in (1, 3):
            assert all(torch.equal(a, b) for a, b in zip(outs[4], outs[path])), f"baked vs lut.path={path}"
        # anchor: 32 rows of one noisy frame against the oracle
        rows = 32
        crop = src[0][: W4K * rows * 4].cpu().numpy()
        want = orc.colorlut(orc.Lut(text=text), crop, W4K, rows)
        assert np.array_equal(outs[4][0][: W4K * rows * 4].cpu().numpy(), want)


def test_hsv_tables_and_compute_kernels_agree_on_64_4k_frames(orc):
    with _ctx() as ctx:
        src = _batch(64, W4K, H4K)
        fp, dp = g.HsvFilterParams(*util.CFG2), g.HsvDetectorParams(*util.DET_CFG4)
        res = {}
        for path in (1, 2):   # per-pixel kernels, function table
            ctx.set_option("hsv.path", path)
            filt = [t.clone() for t in src]
            ctx.hsvfilter_batch(_frames(filt, W4K, H4K, "RGBA"), fp)
            det = [torch.zeros_like(t) for t in src]
            ctx.hsvdetector_batch(_frames(src, W4K, H4K, "BGRx"), _frames(det, W4K, H4K, "RGBA"), dp)
            ctx.synchronize()
            res[path] = (filt, det)
        assert all(torch.equal(a, b) for a, b in zip(res[1][0], res[2][0])), "hsvfilter table vs kernel"
        assert all(torch.equal(a, b) for a, b in zip(res[1][1], res[2][1])), "hsvdetector table vs kernel"
        # hsvdetector BGRx -> RGBA: colour bytes carried over (swapped), alpha is a mask
        for s, d in zip(src[:8], res[2][1][:8]):
            s4, d4 = s.view(-1, 4), d.view(-1, 4)
            assert torch.equal(d4[:, 0], s4[:, 2]) and torch.equal(d4[:, 1], s4[:, 1]) and torch.equal(d4[:, 2], s4[:, 0])
            a = d4[:, 3]
            assert bool(((a == 0) | (a == 255)).all())
        rows = 16
        crop = src[0][: W4K * rows * 4].cpu().numpy()
        assert np.array_equal(res[2][0][0][: W4K * rows * 4].cpu().numpy(), orc.hsvfilter(crop, W4K, rows, "RGBA", util.CFG2))
        assert np.array_equal(res[2][1][0][: W4K * rows * 4].cpu().numpy(),
                              orc.hsvdetector(crop, W4K, rows, "BGRx", "RGBA", util.DET_CFG4))


def test_identity_hsvfilter_moves_bytes_by_at_most_one_code():
    with _ctx() as ctx:
        src = _batch(16, W4K, H4K)
        out = [t.clone() for t in src]
        ctx.hsvfilter_batch(_frames(out, W4K, H4K, "RGBA"), g.HsvFilterParams(*util.IDENTITY))
        ctx.synchronize()
        for s, o in zip(src, out):
            d = (s.view(-1, 4).to(torch.int16) - o.view(-1, 4).to(torch.int16)).abs()
            assert int(d[:, :3].max()) <= 1 and int(d[:, 3].max()) == 0


def test_chain_equals_two_passes_on_8k_frames(orc):
    """cfg5: 8K frames through colorlut ! hsvfilter — fused pass (function table and per-pixel
    kernel) against the two element passes, 16 frames; one crop against the oracle."""
    text = frames.cube_text_3d(33)
    with _ctx() as ctx:
        ctx.set_lut_from_cube(g.parse_cube(text))
        src = _batch(16, W8K, H8K, classes=("noise", "grad"))
        fp = g.HsvFilterParams(*util.CFG2)
        two = [torch.zeros_like(t) for t in src]
        ctx.colorlut_batch(_frames(src, W8K, H8K, "RGBA"), _frames(two, W8K, H8K, "RGBA"))
        ctx.hsvfilter_batch(_frames(two, W8K, H8K, "RGBA"), fp)
        ctx.synchronize()
        for path in (1, 2):
            ctx.set_option("hsv.path", path)
            fused = [torch.zeros_like(t) for t in src]
            ctx.chain_lut_hsv_batch(_frames(src, W8K, H8K, "RGBA"), _frames(fused, W8K, H8K, "RGBA"), fp)
            ctx.synchronize()
            assert all(torch.equal(a, b) for a, b in zip(two, fused)), f"hsv.path={path}"
        rows = 8
        crop = src[0][: W8K * rows * 4].cpu().numpy()
        want = orc.hsvfilter(orc.colorlut(orc.Lut(text=text), crop, W8K, rows), W8K, rows, "RGBA", util.CFG2)
        assert np.array_equal(two[0][: W8K * rows * 4].cpu().numpy(), want)


def test_rgba64_op_and_direct_kernel_agree_on_4k_frames(orc):
    """cfg3: RGBA64 through a 33^3 LUT — the packed-pair delta-table op against the direct 8-corner
    kernel, 16 frames of every content class, both byte orders; one crop against the oracle."""
    text = frames.cube_text_3d(33)
    with _ctx() as ctx:
        ctx.set_lut_from_cube(g.parse_cube(text))
        for fmt, dt in (("RGBA64_LE", "<u2"), ("RGBA64_BE", ">u2")):
            base8 = [frames.frame_of_class(c, W4K, H4K, i).reshape(-1) for i, c in enumerate(("noise", "grad", "bars", "rand"))]
            # 8-bit codes spread over the 16-bit range, plus low-order noise so that fractions vary
            src = []
            for i in range(16):
                v = base8[i % 4].astype(np.uint32) * 257 + ((np.arange(base8[0].size, dtype=np.uint32) * 2654435761 + i) >> 27) % 200
                src.append(torch.from_numpy(np.minimum(v, 65535).astype(dt).view(np.uint8)).cuda())
            outs = {}
            for path in (4, 1):
                ctx.set_option("lut.path", path)
                dst = [torch.zeros_like(t) for t in src]
                ctx.colorlut_batch(_frames(src, W4K, H4K, fmt), _frames(dst, W4K, H4K, fmt))
                ctx.synchronize()
                assert ctx.get_option("lut.path_active") == (7 if path == 4 else 0)
                outs[path] = dst
            assert all(torch.equal(a, b) for a, b in zip(outs[4], outs[1])), fmt
            rows = 8
            crop = src[0][: W4K * rows * 8].cpu().numpy()
            want = orc.colorlut(orc.Lut(text=text), crop, W4K, rows, fmt)
            assert np.array_equal(outs[4][0][: W4K * rows * 8].cpu().numpy(), want), fmt
