"""\"host.register\": recurring pageable frames are page-locked in place on second sight; results are
the oracle's on every path; no registration survives host_memory_released / ctx_destroy."""
import ctypes

import numpy as np
import pytest

import util
import gst_plugins_rs_b200 as g
from gst_plugins_rs_b200 import frames
from gst_plugins_rs_b200.api import frame_of, host_is_pinned

pytestmark = pytest.mark.gpu

W, H = 1024, 96   # 393,216 bytes per frame: above the 64 KiB registration threshold


def _addr(a):
    return a.ctypes.data


def test_second_sight_registers_and_ctx_destroy_unregisters(orc):
    src = frames.frame_rand(W, H, 4, 1).reshape(-1).copy()
    dst = np.zeros_like(src)
    want = orc.hsvdetector(src, W, H, "BGRx", "RGBA", util.DET_CFG4)
    ctx = g.Context(0)
    ctx.set_option("host.register", 1)
    fin, fout = frame_of(src, W, H, "BGRx"), frame_of(dst, W, H, "RGBA")
    for sight in range(4):
        dst[:] = 0
        ctx.hsvdetector(fin, fout, g.HsvDetectorParams(*util.DET_CFG4))
        assert np.array_equal(dst, want), sight
        pinned = host_is_pinned(_addr(src)) and host_is_pinned(_addr(dst))
        assert pinned == (sight >= 1), sight     # first sight bounces, then registered in place
    assert ctx.get_option("host.registered_bytes") == 2 * W * H * 4
    ctx.close()                                   # stop: nothing stays registered
    assert not host_is_pinned(_addr(src)) and not host_is_pinned(_addr(dst))


def test_released_memory_is_forgotten_and_inplace_frames_register_once(orc):
    buf = frames.frame_rand(W, H, 4, 2).reshape(-1).copy()
    with g.Context(0) as ctx:
        ctx.set_option("host.register", 1)
        f = frame_of(buf, W, H, "RGBA")
        cur = buf.copy()
        for _ in range(3):
            cur = orc.hsvfilter(cur, W, H, "RGBA", util.CFG2)
            ctx.hsvfilter(f, g.HsvFilterParams(*util.CFG2))
            assert np.array_equal(buf, cur)
        assert host_is_pinned(_addr(buf))
        assert ctx.get_option("host.registered_bytes") == W * H * 4
        ctx.host_memory_released(_addr(buf) + 4096)     # any address inside the range
        assert not host_is_pinned(_addr(buf))
        assert ctx.get_option("host.registered_bytes") == 0
        ctx.hsvfilter(f, g.HsvFilterParams(*util.CFG2))  # met again: first sight of a fresh history
        assert not host_is_pinned(_addr(buf))
        ctx.set_option("host.register", 0)               # switching off drops everything
        ctx.hsvfilter(f, g.HsvFilterParams(*util.CFG2))
        assert not host_is_pinned(_addr(buf))


def test_budget_evicts_least_recently_used(orc):
    bufs = [frames.frame_rand(W, H, 4, 10 + i).reshape(-1).copy() for i in range(3)]
    with g.Context(0) as ctx:
        ctx.set_option("host.register", 1)
        ctx.set_option("host.register_budget", 2 * W * H * 4)   # room for two frames
        p = g.HsvFilterParams(*util.IDENTITY)
        for rounds in range(2):
            for b in bufs[:2]:
                ctx.hsvfilter(frame_of(b, W, H, "RGBA"), p)
        assert [host_is_pinned(_addr(b)) for b in bufs] == [True, True, False]
        for _ in range(2):
            ctx.hsvfilter(frame_of(bufs[2], W, H, "RGBA"), p)    # third buffer: evicts the LRU (bufs[0])
        assert [host_is_pinned(_addr(b)) for b in bufs] == [False, True, True]
        assert ctx.get_option("host.registered_bytes") == 2 * W * H * 4


def test_small_and_already_pinned_frames_are_left_alone(orc):
    import torch
    with g.Context(0) as ctx:
        ctx.set_option("host.register", 1)
        small = frames.frame_rand(64, 8, 4, 3).reshape(-1).copy()       # 2 KiB: not worth a registration
        for _ in range(3):
            ctx.hsvfilter(frame_of(small, 64, 8, "RGBA"), g.HsvFilterParams(*util.CFG2))
        assert not host_is_pinned(_addr(small))
        t = torch.from_numpy(frames.frame_rand(W, H, 4, 4).reshape(-1).copy()).pin_memory()
        for _ in range(3):
            ctx.hsvfilter(frame_of(t, W, H, "RGBA"), g.HsvFilterParams(*util.CFG2))
        assert ctx.get_option("host.registered_bytes") == 0
    assert host_is_pinned(t.data_ptr())     # torch's own pinning is untouched by ctx_destroy
