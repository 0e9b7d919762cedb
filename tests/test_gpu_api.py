"""C-ABI behaviour beyond per-pixel parity: argument validation, batching, streams, empty and
very large frames, in == out, stats (include/b200vf.h conventions)."""
import ctypes as C

import numpy as np
import pytest

import util
from gst_plugins_rs_b200 import frames
from gst_plugins_rs_b200.api import (B200VFError, ERR_INVALID_ARG, ERR_UNSUPPORTED_FORMAT, Frame,
                                     frame_of)

pytestmark = pytest.mark.gpu


def test_argument_validation(ctx, vf):
    import torch
    t = torch.zeros(64 * 4, dtype=torch.uint8, device="cuda")
    p = vf.HsvFilterParams(*util.CFG2)
    with pytest.raises(B200VFError) as e:   # stride smaller than a row
        ctx.hsvfilter(frame_of(t, 64, 1, "RGBA", stride=100), p)
    assert e.value.status == ERR_INVALID_ARG
    with pytest.raises(B200VFError) as e:   # NULL data
        ctx.hsvfilter(Frame(None, 256, 64, 1, 0, 1), p)
    assert e.value.status == ERR_INVALID_ARG
    with pytest.raises(B200VFError) as e:   # format outside hsvfilter's caps
        ctx.hsvfilter(frame_of(t, 32, 1, "RGBA64_LE"), p)
    assert e.value.status == ERR_UNSUPPORTED_FORMAT
    with pytest.raises(B200VFError) as e:   # unknown format id
        ctx.hsvfilter(Frame(t.data_ptr(), 256, 64, 1, 99, 1), p)
    assert e.value.status == ERR_UNSUPPORTED_FORMAT
    d = vf.HsvDetectorParams(*util.DET_CFG4)
    with pytest.raises(B200VFError) as e:   # detector: RGBA is not a sink format
        ctx.hsvdetector(frame_of(t, 64, 1, "RGBA"), frame_of(t, 64, 1, "RGBA"), d)
    assert e.value.status == ERR_UNSUPPORTED_FORMAT
    with pytest.raises(B200VFError) as e:   # in/out size mismatch
        ctx.hsvdetector(frame_of(t, 64, 1, "BGRx"), frame_of(t, 32, 2, "RGBA"), d)
    assert e.value.status == ERR_INVALID_ARG
    with pytest.raises(B200VFError) as e:   # mixed memory kinds in one call
        ctx.hsvdetector(frame_of(t, 64, 1, "BGRx"), frame_of(np.zeros(256, np.uint8), 64, 1, "RGBA"), d)
    assert e.value.status == ERR_INVALID_ARG
    with pytest.raises(B200VFError):
        ctx.set_option("no.such.option", 1)
    assert ctx.get_option("lut.path") == 0


def test_empty_frames_are_noops(ctx, vf):
    import torch
    t = torch.full((16,), 7, dtype=torch.uint8, device="cuda")
    ctx.hsvfilter(frame_of(t, 0, 4, "RGBA", stride=4), vf.HsvFilterParams(*util.CFG2))
    ctx.hsvfilter(frame_of(t, 4, 0, "RGBA"), vf.HsvFilterParams(*util.CFG2))
    ctx.hsvfilter_batch([], vf.HsvFilterParams(*util.CFG2))
    ctx.synchronize()
    assert (t.cpu().numpy() == 7).all()


def test_tiny_and_ragged_frames(ctx, orc):
    """1x1 … 9x3 frames, every width class of the vector path (W*H mod 4, tails, single rows)."""
    for w in (1, 2, 3, 4, 5, 7, 8, 9, 31, 33, 1023, 1025):
        for h in (1, 2, 3):
            src = frames.random_bytes(w * h * 4, w * 10 + h)
            got = util.gpu_hsvfilter(ctx, src, w, h, "BGRA", util.CFG2)
            assert np.array_equal(got, orc.hsvfilter(src, w, h, "BGRA", util.CFG2)), (w, h)
            gd = util.gpu_hsvdetector(ctx, src, w, h, "xRGB", "ABGR", util.DET_CFG4)
            assert np.array_equal(gd, orc.hsvdetector(src, w, h, "xRGB", "ABGR", util.DET_CFG4)), (w, h)
            rgb = frames.random_bytes(((w * 3 + 3) & ~3) * h, w + h)
            g3 = util.gpu_hsvfilter(ctx, rgb, w, h, "RGB", util.CFG2, stride=(w * 3 + 3) & ~3)
            assert np.array_equal(g3, orc.hsvfilter(rgb, w, h, "RGB", util.CFG2, stride=(w * 3 + 3) & ~3))


def test_8k_frame_and_large_batch(ctx, orc, vf):
    """cfg5 frame size (7680x4320) and a batch larger than one launch holds (> 64 frames)."""
    import torch
    w, h = 7680, 4320
    src = frames.frame_grad(w, h)
    got = util.gpu_hsvfilter(ctx, src, w, h, "RGBA", util.CFG2)
    assert np.array_equal(got, orc.hsvfilter(src, w, h, "RGBA", util.CFG2))
    # 70 small frames, mixed geometry in one batch call
    ws = [(320 + 4 * (i % 3), 48 + (i % 2)) for i in range(70)]
    srcs = [frames.random_bytes(a * b * 4, i) for i, (a, b) in enumerate(ws)]
    ts = [torch.from_numpy(s.copy()).cuda() for s in srcs]
    ctx.reset_stats()
    ctx.hsvfilter_batch([frame_of(t, a, b, "RGBA") for t, (a, b) in zip(ts, ws)],
                        vf.HsvFilterParams(*util.CFG2))
    ctx.synchronize()
    for t, s, (a, b) in zip(ts, srcs, ws):
        assert np.array_equal(t.cpu().numpy(), orc.hsvfilter(s, a, b, "RGBA", util.CFG2))
    st = ctx.stats()
    assert st["frames"] == 70 and st["kernel_launches"] >= 2


def test_colorlut_in_place_and_external_stream(ctx, orc, vf):
    """in == out works although the element is NeverInPlace; device work follows a caller stream."""
    import torch
    text = frames.cube_text_3d(17)
    ctx.set_lut_from_cube(vf.parse_cube(text))
    lut = orc.Lut(text=text)
    w, h = 1000, 37
    src = frames.frame_rand(w, h, 4, 3).reshape(-1)
    s = torch.cuda.Stream()
    ctx.set_stream(s.cuda_stream)
    assert ctx.get_stream() == s.cuda_stream
    with torch.cuda.stream(s):
        t = torch.from_numpy(src.copy()).cuda(non_blocking=False)
        f = frame_of(t, w, h, "RGBA")
        ctx.colorlut(f, f)
        out = t.clone()          # ordered after the kernel on the same stream
    s.synchronize()
    assert np.array_equal(out.cpu().numpy(), orc.colorlut(lut, src, w, h))


def test_host_pipeline_chunking_and_pageable(ctx, orc, vf):
    """System-memory frames: tiny chunks (many pipeline slots reused), pageable and pinned,
    padded strides; only row bytes may be written."""
    import torch
    ctx.set_option("host.chunk_bytes", 64 * 1024)
    w, h, pad = 1921, 270, 28
    stride = w * 4 + pad
    src = frames.random_bytes(stride * h, 5)
    want = src.copy()
    want[:] = orc.hsvfilter(src, w, h, "RGBx", util.CFG2, stride=stride)
    for pinned in (False, True):
        buf = torch.from_numpy(src.copy())
        buf = buf.pin_memory() if pinned else buf
        ctx.hsvfilter(frame_of(buf, w, h, "RGBx", stride), vf.HsvFilterParams(*util.CFG2))
        assert np.array_equal(buf.numpy(), want), f"pinned={pinned}"
    st = ctx.stats()
    assert st["h2d_bytes"] == 2 * w * h * 4 and st["d2h_bytes"] == 2 * w * h * 4


def test_row_band_split_equals_whole_frame(ctx, orc, vf):
    """A frame processed as contiguous row bands (how one huge frame would be split over GPUs,
    SURVEY.md §8e) gives the same bytes as one call: frames are plain (pointer, stride) views."""
    import torch
    from gst_plugins_rs_b200 import sharding
    w, h = 1279, 203
    src = frames.frame_rand(w, h, 4, 21).reshape(-1)
    whole = orc.hsvfilter(src, w, h, "RGBA", util.CFG2)
    t = torch.from_numpy(src.copy()).cuda()
    for first, n in sharding.row_bands(h, 8):
        ctx.hsvfilter(frame_of(t, w, n, "RGBA", w * 4, offset=first * w * 4),
                      vf.HsvFilterParams(*util.CFG2))
    ctx.synchronize()
    assert np.array_equal(t.cpu().numpy(), whole)


def test_contexts_are_independent_across_threads(orc, vf):
    """One context per element instance, each driven from its own streaming thread:
    four threads, four contexts, different elements and settings, host and device frames at once."""
    import threading
    import torch
    w, h = 1024, 200
    src = frames.frame_rand(w, h, 4, 31).reshape(-1)
    text = frames.cube_text_3d(17)
    lut = orc.Lut(text=text)
    jobs = {
        "hsv_dev": lambda: orc.hsvfilter(src, w, h, "RGBA", util.CFG2),
        "hsv_host": lambda: orc.hsvfilter(src, w, h, "BGRA", (-90.0, 0.5, 0.2, 1.5, -0.1)),
        "lut_dev": lambda: orc.colorlut(lut, src, w, h),
        "det_host": lambda: orc.hsvdetector(src, w, h, "BGRx", "RGBA", util.DET_CFG4),
    }
    want = {k: f() for k, f in jobs.items()}
    errors = []

    def worker(kind):
        try:
            with vf.Context(0) as c:
                for _ in range(25):
                    if kind == "hsv_dev":
                        t = torch.from_numpy(src.copy()).cuda()
                        c.hsvfilter(frame_of(t, w, h, "RGBA"), vf.HsvFilterParams(*util.CFG2))
                        c.synchronize()
                        got = t.cpu().numpy()
                    elif kind == "hsv_host":
                        got = src.copy()
                        c.hsvfilter(frame_of(got, w, h, "BGRA"),
                                    vf.HsvFilterParams(-90.0, 0.5, 0.2, 1.5, -0.1))
                    elif kind == "lut_dev":
                        c.set_lut_from_cube(vf.parse_cube(text))
                        s = torch.from_numpy(src.copy()).cuda()
                        d = torch.empty_like(s)
                        c.colorlut(frame_of(s, w, h, "RGBA"), frame_of(d, w, h, "RGBA"))
                        c.synchronize()
                        got = d.cpu().numpy()
                    else:
                        got = np.zeros(w * h * 4, np.uint8)
                        mine = src.copy()  # keep the buffer alive: a frame only borrows the pointer
                        c.hsvdetector(frame_of(mine, w, h, "BGRx"), frame_of(got, w, h, "RGBA"),
                                      vf.HsvDetectorParams(*util.DET_CFG4))
                    if not np.array_equal(got, want[kind]):
                        errors.append(kind)
                        return
        except Exception as e:  # noqa: BLE001
            errors.append(f"{kind}: {e!r}")

    threads = [threading.Thread(target=worker, args=(k,)) for k in jobs]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


def test_two_contexts_ordered_with_wait_for(orc):
    """Chained device-memory elements keep their own streams; the frame hand-over is ordered on the
    device by b200vf_ctx_wait_for (no shared stream handle, no host sync)."""
    import torch
    import util
    import gst_plugins_rs_b200 as g
    from gst_plugins_rs_b200 import frames
    from gst_plugins_rs_b200.api import B200VFError, frame_of
    w, h = 1920, 1080
    text = frames.cube_text_3d(17)
    lut = orc.Lut(text=text)
    with g.Context(0) as a, g.Context(0) as b:
        a.set_lut_from_cube(g.parse_cube(text))
        assert a.get_stream() != b.get_stream()
        for i in range(4):
            src = frames.frame_of_class(("noise", "rand")[i % 2], w, h, i).reshape(-1)
            s = torch.from_numpy(src.copy()).cuda()
            mid = torch.zeros_like(s)
            a.colorlut(frame_of(s, w, h, "RGBA"), frame_of(mid, w, h, "RGBA"))
            b.wait_for(a)
            b.hsvfilter(frame_of(mid, w, h, "RGBA"), g.HsvFilterParams(*util.CFG2))
            b.synchronize()
            want = orc.hsvfilter(orc.colorlut(lut, src, w, h), w, h, "RGBA", util.CFG2)
            assert np.array_equal(mid.cpu().numpy(), want), i
        b.wait_for(b)   # a no-op
        with pytest.raises(B200VFError):
            b.lib.b200vf_ctx_wait_for.restype
            b._check(b.lib.b200vf_ctx_wait_for(b.h, None))
