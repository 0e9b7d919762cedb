"""Device-wide table cache, lazily built LUT tables, atomic set_lut and the measured path policy
(vf_tables.cpp, vf_abi.cpp): contexts that need the same function share one 64 MiB table, tables are
built by the path that needs them, and whatever kernel serves, the bytes are the oracle's."""
import numpy as np
import pytest

import util
import gst_plugins_rs_b200 as g
from gst_plugins_rs_b200 import frames
from gst_plugins_rs_b200.api import B200VFError

pytestmark = pytest.mark.gpu

TABLE = 64 << 20


def _run_filter(ctx, orc, settings, seed=1, w=640, h=48):
    src = frames.frame_rand(w, h, 4, seed).reshape(-1)
    got = util.gpu_hsvfilter(ctx, src, w, h, "RGBA", settings)
    assert np.array_equal(got, orc.hsvfilter(src.copy(), w, h, "RGBA", settings))


def test_function_tables_are_shared_between_contexts(orc):
    with g.Context(0) as a, g.Context(0) as b, g.Context(0) as c:
        base = a.get_option("tables.device_count")
        for ctx in (a, b, c):
            ctx.set_option("hsv.path", 2)
        _run_filter(a, orc, util.CFG2)
        assert a.get_option("tables.device_count") == base + 1
        _run_filter(b, orc, util.CFG2, seed=2)          # same function: same table
        assert a.get_option("tables.device_count") == base + 1
        _run_filter(c, orc, util.IDENTITY, seed=3)      # another function: its own table
        assert a.get_option("tables.device_count") == base + 2
        assert a.get_option("tables.device_bytes") == (base + 2) * TABLE
        _run_filter(b, orc, util.IDENTITY, seed=4)      # b moves over to c's table, a keeps CFG2's
        assert a.get_option("tables.device_count") == base + 2
        _run_filter(a, orc, util.IDENTITY, seed=5)      # last user of CFG2's table leaves: freed
        assert a.get_option("tables.device_count") == base + 1
    with g.Context(0) as d:
        assert d.get_option("tables.device_count") == base  # contexts gone, tables gone


def test_baked_lut_table_is_shared_by_content_not_by_context(orc):
    text33, text17 = frames.cube_text_3d(33), frames.cube_text_3d(17)
    w, h = 256, 32
    src = frames.frame_rand(w, h, 4, 9).reshape(-1)
    with g.Context(0) as a, g.Context(0) as b:
        base = a.get_option("tables.device_count")
        a.set_lut_from_cube(g.parse_cube(text33))
        b.set_lut_from_cube(g.parse_cube(text33))
        assert a.get_option("lut.tables_built") == 0     # nothing is built at set_lut
        for ctx in (a, b):
            got = util.gpu_colorlut(ctx, src, w, h)
            assert np.array_equal(got, orc.colorlut(orc.Lut(text=text33), src, w, h))
            assert ctx.get_option("lut.tables_built") == 4   # only the baked table, no RX / RG
        assert a.get_option("tables.device_count") == base + 1
        b.set_lut_from_cube(g.parse_cube(text17))          # different content: different table
        got = util.gpu_colorlut(b, src, w, h)
        assert np.array_equal(got, orc.colorlut(orc.Lut(text=text17), src, w, h))
        assert a.get_option("tables.device_count") == base + 2
        b.set_option("lut.interpolation", 1)               # same LUT, other interpolation: new key
        util.gpu_colorlut(b, src, w, h)
        assert a.get_option("tables.device_count") == base + 2   # b swapped its table, a keeps 33^3
        a.clear_lut()
        assert b.get_option("tables.device_count") == base + 1


@pytest.mark.parametrize("path,mask", [(1, 0), (2, 1), (3, 3), (4, 4), (0, 4)])
def test_lut_tables_are_built_by_the_path_that_needs_them(ctx, orc, path, mask):
    text = frames.cube_text_3d(9)
    w, h = 320, 24
    src = frames.frame_rand(w, h, 4, 11).reshape(-1)
    ctx.set_lut_from_cube(g.parse_cube(text))
    ctx.set_option("lut.path", path)
    got = util.gpu_colorlut(ctx, src, w, h)
    assert np.array_equal(got, orc.colorlut(orc.Lut(text=text), src, w, h))
    assert ctx.get_option("lut.tables_built") == mask
    # 16-bit frames never need any of them
    src16 = frames.random_bytes(w * h * 8, 12)
    got = util.gpu_colorlut(ctx, src16, w, h, "RGBA64_LE")
    assert np.array_equal(got, orc.colorlut(orc.Lut(text=text), src16, w, h, "RGBA64_LE"))
    assert ctx.get_option("lut.tables_built") == mask


def test_failed_set_lut_keeps_the_previous_lut(ctx, orc):
    text = frames.cube_text_3d(5)
    cube = g.parse_cube(text)
    w, h = 128, 16
    src = frames.frame_rand(w, h, 4, 13).reshape(-1)
    ctx.set_lut_from_cube(cube)
    want = orc.colorlut(orc.Lut(text=text), src, w, h)
    assert np.array_equal(util.gpu_colorlut(ctx, src, w, h), want)
    for kind, size in ((3, 1), (3, 257), (1, 1), (1, 65537), (2, 8)):
        with pytest.raises(B200VFError):
            ctx.set_lut(kind, size, np.zeros(16, np.float32), np.ones(3, np.float32), np.zeros(3, np.float32))
    assert np.array_equal(util.gpu_colorlut(ctx, src, w, h), want)   # still the 5^3 LUT


def test_colorlut_auto_policy_stays_exact_while_it_measures(ctx, orc):
    """auto ("lut.path" = 0) times the baked table against the direct kernel on real frames
    (frames of >= 2^20 pixels) and keeps re-timing; every launch must give the oracle's bytes
    whichever kernel served it, and both kernels must have been seen."""
    text = frames.cube_text_3d(17)
    lut = orc.Lut(text=text)
    ctx.set_lut_from_cube(g.parse_cube(text))
    w, h = 1280, 1024
    seen = set()
    for i in range(12):
        src = frames.frame_of_class(("rand", "noise", "grad")[i % 3], w, h, i).reshape(-1)
        got = util.gpu_colorlut(ctx, src, w, h)
        assert np.array_equal(got, orc.colorlut(lut, src, w, h)), i
        seen.add(ctx.get_option("lut.path_active"))
    assert seen == {0, 4}, seen


def test_concurrent_contexts_build_and_share_one_table(orc):
    """Eight threads, eight contexts, the same function at the same moment (ctypes drops the GIL in
    the calls): one of them builds the table on its stream, the others wait for its event on theirs;
    every result is the oracle's and one table exists."""
    import threading
    w, h = 512, 64
    srcs = [frames.frame_rand(w, h, 4, 300 + i).reshape(-1) for i in range(8)]
    wants = [orc.hsvfilter(s.copy(), w, h, "RGBA", util.CFG2) for s in srcs]
    ctxs = [g.Context(0) for _ in range(8)]
    base = ctxs[0].get_option("tables.device_count")
    for c in ctxs:
        c.set_option("hsv.path", 2)
    start = threading.Barrier(8)
    errors = []

    def work(i):
        try:
            start.wait()
            for _ in range(3):
                got = util.gpu_hsvfilter(ctxs[i], srcs[i], w, h, "RGBA", util.CFG2)
                if not np.array_equal(got, wants[i]):
                    errors.append(i)
        except Exception as e:  # pragma: no cover
            errors.append(repr(e))
    ts = [threading.Thread(target=work, args=(i,)) for i in range(8)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert errors == []
    assert ctxs[0].get_option("tables.device_count") == base + 1
    for c in ctxs:
        c.close()
