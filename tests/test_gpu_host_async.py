"""\"host.async\": calls on page-locked host frames return once queued and are told apart by
tickets; the pixels after b200vf_ctx_host_wait are the oracle's; pageable frames stay synchronous."""
import numpy as np
import pytest
import torch

import util
import gst_plugins_rs_b200 as g
from gst_plugins_rs_b200 import frames
from gst_plugins_rs_b200.api import B200VFError, frame_array, frame_of

pytestmark = pytest.mark.gpu

W, H = 640, 360


def _pinned(arr):
    return torch.from_numpy(np.ascontiguousarray(arr).reshape(-1).copy()).pin_memory()


def test_calls_in_flight_give_the_oracles_pixels(orc):
    text = frames.cube_text_3d(17)
    lut = orc.Lut(text=text)
    n = 40                                        # more calls than tickets (16) and ring slots
    srcs = [frames.frame_rand(W, H, 4, 100 + i).reshape(-1) for i in range(n)]
    wants = [orc.colorlut(lut, s, W, H) for s in srcs]
    with g.Context(0) as ctx:
        ctx.set_lut_from_cube(g.parse_cube(text))
        ctx.set_option("host.chunk_bytes", 64 << 10)   # a dozen chunks per frame: slots are reused within a call
        ctx.set_option("host.async", 1)
        assert ctx.get_option("host.async") == 1
        h_in = [_pinned(s) for s in srcs]
        h_out = [torch.zeros(W * H * 4, dtype=torch.uint8).pin_memory() for _ in srcs]
        base = ctx.host_ticket()
        tickets = []
        for i in range(n):                        # one frame of latency, like an element that queues buffers
            ctx.colorlut(frame_of(h_in[i], W, H, "RGBA"), frame_of(h_out[i], W, H, "RGBA"))
            tickets.append(ctx.host_ticket())
            if i:
                ctx.host_wait(tickets[i - 1])
                assert np.array_equal(h_out[i - 1].numpy(), wants[i - 1]), i - 1
        assert tickets == list(range(base + 1, base + n + 1))
        ctx.host_wait(tickets[-1])
        assert np.array_equal(h_out[-1].numpy(), wants[-1])
        ctx.host_wait(tickets[0])                 # long complete: returns at once
        with pytest.raises(B200VFError):
            ctx.host_wait(tickets[-1] + 1)        # no such call yet


def test_nothing_waited_for_until_synchronize(orc):
    """The caller never waits per call: the library bounds the calls in flight itself, and
    ctx.synchronize() completes them all."""
    n = 24
    srcs = [frames.frame_rand(W, H, 4, 200 + i).reshape(-1) for i in range(n)]
    wants = [orc.hsvdetector(s, W, H, "BGRx", "RGBA", util.DET_CFG4) for s in srcs]
    with g.Context(0) as ctx:
        ctx.set_option("host.async", 1)
        h_in = [_pinned(s) for s in srcs]
        h_out = [torch.zeros(W * H * 4, dtype=torch.uint8).pin_memory() for _ in srcs]
        p = g.HsvDetectorParams(*util.DET_CFG4)
        for i in range(n):
            ctx.hsvdetector(frame_of(h_in[i], W, H, "BGRx"), frame_of(h_out[i], W, H, "RGBA"), p)
        ctx.synchronize()
        for i in range(n):
            assert np.array_equal(h_out[i].numpy(), wants[i]), i


def test_pageable_frames_stay_synchronous_and_batches_mix(orc):
    srcs = [frames.frame_rand(W, H, 4, 300 + i).reshape(-1).copy() for i in range(4)]
    with g.Context(0) as ctx:
        ctx.set_option("host.async", 1)
        p = g.HsvFilterParams(*util.CFG2)
        # in place, pageable: complete on return although the option is on
        buf = srcs[0].copy()
        ctx.hsvfilter(frame_of(buf, W, H, "RGBA"), p)
        assert np.array_equal(buf, orc.hsvfilter(srcs[0], W, H, "RGBA", util.CFG2))
        t_sync = ctx.host_ticket()
        ctx.host_wait(t_sync)
        # a pinned call in flight, then a batch holding a pageable frame: the batch's return
        # completes both (the copy-out stream runs in order)
        pin = _pinned(srcs[1])
        ctx.hsvfilter(frame_of(pin, W, H, "RGBA"), p)
        pin2, page = _pinned(srcs[2]), srcs[3].copy()
        ctx.hsvfilter_batch(frame_array([frame_of(pin2, W, H, "RGBA"), frame_of(page, W, H, "RGBA")]), p)
        for got, src in ((pin.numpy(), srcs[1]), (pin2.numpy(), srcs[2]), (page, srcs[3])):
            assert np.array_equal(got, orc.hsvfilter(src, W, H, "RGBA", util.CFG2))
        assert ctx.host_ticket() == t_sync + 2
        # switching the option off completes what is in flight
        pin3 = _pinned(srcs[0])
        ctx.hsvfilter(frame_of(pin3, W, H, "RGBA"), p)
        ctx.set_option("host.async", 0)
        assert np.array_equal(pin3.numpy(), orc.hsvfilter(srcs[0], W, H, "RGBA", util.CFG2))
        # and synchronous calls get tickets too
        ctx.hsvfilter(frame_of(pin3, W, H, "RGBA"), p)
        assert ctx.host_ticket() == t_sync + 4
        ctx.host_wait(ctx.host_ticket())


def test_lut_change_between_calls_in_flight(orc):
    """set_lut while earlier calls are still in flight: each call sees the LUT of its own time."""
    t1 = frames.cube_text_3d(9)
    t2 = frames.cube_text_3d(17, values=1.0 - np.asarray(frames.synthetic_lut_values(17)))
    src = frames.frame_rand(W, H, 4, 400).reshape(-1)
    with g.Context(0) as ctx:
        ctx.set_option("host.async", 1)
        a, b = _pinned(src), _pinned(src)
        oa = torch.zeros(W * H * 4, dtype=torch.uint8).pin_memory()
        ob = torch.zeros(W * H * 4, dtype=torch.uint8).pin_memory()
        ctx.set_lut_from_cube(g.parse_cube(t1))
        ctx.colorlut(frame_of(a, W, H, "RGBA"), frame_of(oa, W, H, "RGBA"))
        ctx.set_lut_from_cube(g.parse_cube(t2))
        ctx.colorlut(frame_of(b, W, H, "RGBA"), frame_of(ob, W, H, "RGBA"))
        ctx.host_wait(ctx.host_ticket())
        assert np.array_equal(oa.numpy(), orc.colorlut(orc.Lut(text=t1), src, W, H))
        assert np.array_equal(ob.numpy(), orc.colorlut(orc.Lut(text=t2), src, W, H))


def test_random_mix_of_geometries_memory_kinds_and_modes(orc):
    """Soak: 60 calls on one context — random element, format, odd widths, padded strides, pinned
    or pageable frames, chunk sizes from a few rows to whole frames, "host.async" toggled at random,
    up to three calls in flight.  Every output equals the oracle's; padding bytes are untouched."""
    rng = np.random.default_rng(2024)
    text = frames.cube_text_3d(9)
    lut = orc.Lut(text=text)
    with g.Context(0) as ctx:
        ctx.set_lut_from_cube(g.parse_cube(text))
        pending = []   # (ticket, got buffer, wanted bytes, input buffer: the library's until waited for)

        def settle(keep):
            while len(pending) > keep:
                t, got, want, _src = pending.pop(0)
                ctx.host_wait(t)
                arr = got.numpy() if hasattr(got, "numpy") else got
                assert np.array_equal(arr, want)

        for it in range(60):
            if rng.integers(4) == 0:
                settle(0)
                ctx.set_option("host.async", int(rng.integers(2)))
            ctx.set_option("host.chunk_bytes", int(rng.choice([0, 4096, 20000, 1 << 20])))
            elem = rng.choice(["colorlut", "hsvfilter", "hsvdetector"])
            w, h = int(rng.integers(1, 700)), int(rng.integers(1, 90))
            pin = [bool(rng.integers(2)), bool(rng.integers(2))]

            def buf(arr, pinned):
                return torch.from_numpy(arr.copy()).pin_memory() if pinned else arr.copy()

            if elem == "colorlut":
                fmt = str(rng.choice(["RGBA", "RGBA64_LE", "RGBA64_BE"]))
                bpp = 4 if fmt == "RGBA" else 8
                s_in, s_out = w * bpp + 8 * int(rng.integers(3)), w * bpp + 8 * int(rng.integers(3))
                src = frames.random_bytes(s_in * h, 1000 + it)
                dst0 = frames.random_bytes(s_out * h, 2000 + it)
                want = dst0.copy()
                orc.colorlut(lut, src, w, h, fmt, src_stride=s_in, dst_stride=s_out, dst=want)
                a, b = buf(src, pin[0]), buf(dst0, pin[1])
                ctx.colorlut(frame_of(a, w, h, fmt, s_in), frame_of(b, w, h, fmt, s_out))
            elif elem == "hsvfilter":
                fmt = str(rng.choice(["RGBx", "BGRA", "xRGB", "RGB", "BGR"]))
                bpp = 3 if fmt in ("RGB", "BGR") else 4
                s_in = (w * bpp + 3) // 4 * 4 + 4 * int(rng.integers(3))
                src = frames.random_bytes(s_in * h, 3000 + it)
                want = src.copy()
                want[:] = orc.hsvfilter(src, w, h, fmt, util.CFG2, stride=s_in)
                a = b = buf(src, pin[0])
                ctx.hsvfilter(frame_of(b, w, h, fmt, s_in), g.HsvFilterParams(*util.CFG2))
            else:
                in_fmt, out_fmt = str(rng.choice(["RGBx", "BGRx", "RGB"])), str(rng.choice(["RGBA", "ABGR"]))
                bpp = 3 if in_fmt == "RGB" else 4
                s_in = (w * bpp + 3) // 4 * 4 + 4 * int(rng.integers(3))
                s_out = w * 4 + 4 * int(rng.integers(3))
                src = frames.random_bytes(s_in * h, 4000 + it)
                dst0 = frames.random_bytes(s_out * h, 5000 + it)
                want = dst0.copy()
                orc.hsvdetector(src, w, h, in_fmt, out_fmt, util.DET_CFG4, in_stride=s_in, out_stride=s_out, dst=want)
                a, b = buf(src, pin[0]), buf(dst0, pin[1])
                ctx.hsvdetector(frame_of(a, w, h, in_fmt, s_in), frame_of(b, w, h, out_fmt, s_out),
                                g.HsvDetectorParams(*util.DET_CFG4))
            pending.append((ctx.host_ticket(), b, want, a))
            settle(int(rng.integers(4)))
        settle(0)
