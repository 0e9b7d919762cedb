"""colorlut with the videoconvert steps folded in (b200vf_colorlut_convert_process_batch): any of
the ten 8-bit packed layouts in, any out.  Expected bytes = the oracle's colorlut on the RGBA view
of the input, re-packed by the byte rules of SURVEY.md Appendix C (colour order / offset per format;
alpha = the source's, 255 when it has none; an x byte receives the same value).  The conversion
rules themselves are GStreamer core's and therefore parity-unpinned; the colour values are pinned."""
import numpy as np
import pytest

import util
import gst_plugins_rs_b200 as g
from gst_plugins_rs_b200 import frames
from gst_plugins_rs_b200.api import BYTES_PER_PIXEL, B200VFError, frame_of

pytestmark = pytest.mark.gpu

# format -> (bpp, r, g, b, alpha-or-padding offset, has real alpha)
LAYOUT = {"RGBA": (4, 0, 1, 2, 3, True), "RGBx": (4, 0, 1, 2, 3, False), "xRGB": (4, 1, 2, 3, 0, False),
          "ARGB": (4, 1, 2, 3, 0, True), "BGRx": (4, 2, 1, 0, 3, False), "BGRA": (4, 2, 1, 0, 3, True),
          "xBGR": (4, 3, 2, 1, 0, False), "ABGR": (4, 3, 2, 1, 0, True), "RGB": (3, 0, 1, 2, None, False),
          "BGR": (3, 2, 1, 0, None, False)}
FORMATS = list(LAYOUT)


def expected(orc, lut, src, w, h, fin, fout, in_stride, out_stride, fill):
    bpp, r, gg, b, a, has_a = LAYOUT[fin]
    rows = np.frombuffer(src, np.uint8).reshape(h, in_stride)[:, : w * bpp].reshape(h, w, bpp)
    rgba = np.empty((h, w, 4), np.uint8)
    rgba[..., 0], rgba[..., 1], rgba[..., 2] = rows[..., r], rows[..., gg], rows[..., b]
    rgba[..., 3] = rows[..., a] if has_a else 255
    res = orc.colorlut(lut, rgba.reshape(-1), w, h).reshape(h, w, 4)
    obpp, orr, og, ob, oa, _ = LAYOUT[fout]
    out = np.full((h, out_stride), fill, np.uint8)
    px = np.empty((h, w, obpp), np.uint8)
    px[..., orr], px[..., og], px[..., ob] = res[..., 0], res[..., 1], res[..., 2]
    if oa is not None:
        px[..., oa] = res[..., 3]
    out[:, : w * obpp] = px.reshape(h, w * obpp)
    return out.reshape(-1)


def run(ctx, src, w, h, fin, fout, in_stride, out_stride, memory, fill=0xA5):
    sbuf = util._buffers(src, memory)
    dbuf = util._buffers(np.full(h * out_stride, fill, np.uint8), memory)
    ctx.colorlut_convert_batch([frame_of(sbuf, w, h, fin, in_stride)], [frame_of(dbuf, w, h, fout, out_stride)])
    ctx.synchronize()
    return util._to_numpy(dbuf).reshape(-1)


@pytest.mark.parametrize("fin", FORMATS)
def test_every_format_pair(ctx, orc, fin):
    text = frames.cube_text_3d(9)
    lut = orc.Lut(text=text)
    ctx.set_lut_from_cube(g.parse_cube(text))
    for k, fout in enumerate(FORMATS):
        for (w, h, pad_in, pad_out, memory) in ((320, 18, 0, 0, "device"), (333, 9, 16, 32, "device"),
                                                (131, 5, 3, 5, "device"), (200, 7, 0, 0, "host")):
            in_stride = w * LAYOUT[fin][0] + pad_in
            out_stride = w * LAYOUT[fout][0] + pad_out
            src = frames.random_bytes(in_stride * h, 7 * k + w)
            got = run(ctx, src, w, h, fin, fout, in_stride, out_stride, memory)
            want = expected(orc, lut, src, w, h, fin, fout, in_stride, out_stride, 0xA5)
            assert np.array_equal(got, want), (fin, fout, w, h, pad_in, pad_out, memory)


def test_convert_with_1d_lut_and_interpolation_modes(ctx, orc):
    w, h = 256, 16
    src = frames.random_bytes(w * h * 4, 5)
    text = frames.cube_text_1d(33)
    ctx.set_lut_from_cube(g.parse_cube(text))
    got = run(ctx, src, w, h, "BGRx", "ARGB", w * 4, w * 4, "device")
    assert np.array_equal(got, expected(orc, orc.Lut(text=text), src, w, h, "BGRx", "ARGB", w * 4, w * 4, 0xA5))
    # "lut.path" does not matter to the conversion path: it always runs from the baked table
    text = frames.cube_text_3d(5)
    ctx.set_lut_from_cube(g.parse_cube(text))
    for path in (1, 3):
        ctx.set_option("lut.path", path)
        got = run(ctx, src, w, h, "ABGR", "RGB", w * 4, w * 3, "device")
        assert np.array_equal(got, expected(orc, orc.Lut(text=text), src, w, h, "ABGR", "RGB", w * 4, w * 3, 0xA5))


def test_convert_rgba_to_rgba_is_the_element_and_errors(ctx, orc):
    w, h = 64, 8
    src = frames.random_bytes(w * h * 4, 1)
    with pytest.raises(B200VFError):   # no LUT yet
        run(ctx, src, w, h, "BGRx", "RGBA", w * 4, w * 4, "device")
    text = frames.cube_text_3d(4)
    ctx.set_lut_from_cube(g.parse_cube(text))
    got = run(ctx, src, w, h, "RGBA", "RGBA", w * 4, w * 4, "device")
    assert np.array_equal(got, orc.colorlut(orc.Lut(text=text), src, w, h))
    src16 = frames.random_bytes(w * h * 8, 2)
    with pytest.raises(B200VFError):   # 16-bit formats are not convertible
        run(ctx, src16, w, h, "RGBA64_LE", "RGBA", w * 8, w * 4, "device")
