"""The tabulated-function path of hsvfilter / hsvdetector / the fused chain ("hsv.path"): the
element's exact compute kernel is run once over all 2^24 colour triples and frames are then served
by one gather per pixel.  It must be indistinguishable from the compute kernels — and therefore
from the oracle — in every format, memory kind and geometry, in forced mode (2) and while auto
mode (0) builds, measures and switches on its own."""
import numpy as np
import pytest

import util
from gst_plugins_rs_b200 import frames
from gst_plugins_rs_b200.api import BYTES_PER_PIXEL

pytestmark = pytest.mark.gpu

FILTER_FORMATS = ["RGBx", "xRGB", "BGRx", "xBGR", "RGBA", "ARGB", "BGRA", "ABGR", "RGB", "BGR"]
DET_IN = ["RGBx", "xRGB", "BGRx", "xBGR", "RGB", "BGR"]
DET_OUT = ["RGBA", "ARGB", "BGRA", "ABGR"]
SETTINGS = [util.CFG2, util.IDENTITY, (-123.25, 0.7, -0.1, 1.3, 0.1), (1234.5, 1.0, 0.25, 1.0, -0.25)]


def _strided(w, h, bpp, seed, pad):
    stride = w * bpp + pad
    buf = frames.random_bytes(stride * h, seed)
    return buf, stride


@pytest.mark.parametrize("fmt", FILTER_FORMATS)
def test_hsvfilter_table_all_formats(ctx, orc, fmt):
    ctx.set_option("hsv.path", 2)
    bpp = BYTES_PER_PIXEL[fmt]
    for k, settings in enumerate(SETTINGS):
        for (w, h, pad, memory) in ((640, 33, 0, "device"), (333, 17, 4 if bpp == 4 else 5, "device"),
                                    (257, 9, 0, "host")):
            src, stride = _strided(w, h, bpp, 10 * k + pad, pad)
            got = util.gpu_hsvfilter(ctx, src, w, h, fmt, settings, stride, memory=memory)
            want = orc.hsvfilter(src.copy(), w, h, fmt, settings, stride=stride)
            assert np.array_equal(got, want), f"{fmt} {settings} {w}x{h}+{pad} {memory}"
            assert ctx.get_option("hsv.table_active") == 1


@pytest.mark.parametrize("in_fmt", DET_IN)
@pytest.mark.parametrize("out_fmt", DET_OUT)
def test_hsvdetector_table_all_format_pairs(ctx, orc, in_fmt, out_fmt):
    ctx.set_option("hsv.path", 2)
    bpp = BYTES_PER_PIXEL[in_fmt]
    for settings in (util.DET_CFG4, util.DET_DEFAULT, (300.0, 180.0, 0.5, 0.5, 0.5, 0.25)):
        for (w, h, pad) in ((512, 20, 0), (301, 7, 8 if bpp == 4 else 3)):
            src, in_stride = _strided(w, h, bpp, 3 + pad, pad)
            out_stride = w * 4 + (16 if pad else 0)
            got = util.gpu_hsvdetector(ctx, src, w, h, in_fmt, out_fmt, settings, in_stride, out_stride)
            want = orc.hsvdetector(src, w, h, in_fmt, out_fmt, settings, in_stride, out_stride,
                                   dst=np.full(h * out_stride, 0xA5, np.uint8))
            assert np.array_equal(got, want), f"{in_fmt}->{out_fmt} {settings} {w}x{h}"
            assert ctx.get_option("hsv.table_active") == 1


def test_tables_exhaustive(ctx, orc):
    """All 2^24 triples through the table, with the pass-through byte set: hsvfilter RGBA / xBGR,
    hsvdetector BGRx->ARGB."""
    ctx.set_option("hsv.path", 2)
    src = frames.all_rgb_frame(0, 1, 2, 3, other_value=77)
    assert np.array_equal(util.gpu_hsvfilter(ctx, src, 4096, 4096, "RGBA", util.CFG2),
                          orc.hsvfilter(src.copy(), 4096, 4096, "RGBA", util.CFG2))
    src = frames.all_rgb_frame(3, 2, 1, 0, other_value=200)   # x,B,G,R
    s = (-17.0, 1.0, 0.0, 1.0, 0.0)
    assert np.array_equal(util.gpu_hsvfilter(ctx, src, 4096, 4096, "xBGR", s),
                          orc.hsvfilter(src.copy(), 4096, 4096, "xBGR", s))
    src = frames.all_rgb_frame(2, 1, 0, 3, other_value=9)     # B,G,R,x
    got = util.gpu_hsvdetector(ctx, src, 4096, 4096, "BGRx", "ARGB", util.DET_CFG4)
    assert np.array_equal(got, orc.hsvdetector(src, 4096, 4096, "BGRx", "ARGB", util.DET_CFG4))


def test_auto_mode_builds_measures_and_follows_settings(ctx, orc):
    """Auto: compute kernels until the settings were stable for 2^25 pixels, then the table is
    built and both ways are timed; every output along the way equals the oracle.  A settings
    change sends it back to the compute kernels immediately."""
    import torch
    import gst_plugins_rs_b200 as g
    from gst_plugins_rs_b200.api import frame_array, frame_of
    assert ctx.get_option("hsv.path") == 0
    w, h, nb = 3840, 2160, 2            # 16.6 Mpixel per call
    src = [frames.frame_of_class(c, w, h, i).reshape(-1) for i, c in enumerate(("grad", "noise"))]
    want = [orc.hsvfilter(s.copy(), w, h, "RGBA", util.CFG2) for s in src]
    seen_table = False
    for step in range(8):
        bufs = [torch.from_numpy(s.copy()).cuda() for s in src]
        ctx.hsvfilter_batch(frame_array([frame_of(t, w, h, "RGBA") for t in bufs]),
                            g.HsvFilterParams(*util.CFG2))
        ctx.synchronize()
        active = ctx.get_option("hsv.table_active")
        if step < 2:
            assert active == 0, "no table before 2^25 stable pixels"
        seen_table |= bool(active)
        for i in range(nb):
            assert np.array_equal(bufs[i].cpu().numpy(), want[i]), f"step {step} frame {i}"
    assert seen_table, "auto mode never tried the table"
    other = (10.0, 1.0, 0.0, 1.0, 0.0)
    bufs = [torch.from_numpy(s.copy()).cuda() for s in src]
    ctx.hsvfilter_batch(frame_array([frame_of(t, w, h, "RGBA") for t in bufs]), g.HsvFilterParams(*other))
    ctx.synchronize()
    assert ctx.get_option("hsv.table_active") == 0
    assert np.array_equal(bufs[0].cpu().numpy(), orc.hsvfilter(src[0].copy(), w, h, "RGBA", other))


def test_chain_table_follows_lut_and_settings(ctx, orc):
    import torch
    import gst_plugins_rs_b200 as g
    from gst_plugins_rs_b200.api import frame_of
    ctx.set_option("hsv.path", 2)
    w, h = 1024, 64
    src = frames.frame_rand(w, h, 4, 5).reshape(-1)

    def run(params):
        s = torch.from_numpy(src.copy()).cuda()
        d = torch.zeros_like(s)
        ctx.chain_lut_hsv_batch([frame_of(s, w, h, "RGBA")], [frame_of(d, w, h, "RGBA")],
                                g.HsvFilterParams(*params))
        ctx.synchronize()
        return d.cpu().numpy()

    for n, interp in ((17, 0), (33, 0), (9, 1)):
        text = frames.cube_text_3d(n)
        ctx.set_lut_from_cube(g.parse_cube(text))
        ctx.set_option("lut.interpolation", interp)
        lut = orc.Lut(text=text)
        mid = orc.colorlut(lut, src, w, h, interpolation="tetrahedral" if interp else "trilinear")
        for params in (util.CFG2, (-75.0, 0.8, 0.1, 1.1, -0.05)):
            want = orc.hsvfilter(mid.copy(), w, h, "RGBA", params)
            assert np.array_equal(run(params), want), f"n={n} interp={interp} {params}"
            assert ctx.get_option("hsv.table_active") == 1


def test_compute_only_mode_never_uses_the_table(ctx, orc):
    ctx.set_option("hsv.path", 1)
    src = frames.all_rgb_frame(0, 1, 2, 3, other_value=1)
    for _ in range(4):
        got = util.gpu_hsvfilter(ctx, src, 4096, 4096, "RGBA", util.CFG2)
        assert ctx.get_option("hsv.table_active") == 0
    assert np.array_equal(got, orc.hsvfilter(src.copy(), 4096, 4096, "RGBA", util.CFG2))
