"""b200vf_group — the in-process frame-parallel dispatcher: frame i of a batch is
processed by member i mod G, results land in out[i], bytes equal the oracle's.  On a 1-GPU box the
members are several contexts on device 0 (round-robin, threads and ordering are exercised the same
way); with >= 2 GPUs the frames really live on different devices."""
import numpy as np
import pytest
import torch

import util
import gst_plugins_rs_b200 as g
from gst_plugins_rs_b200 import frames
from gst_plugins_rs_b200.api import B200VFError, frame_of

pytestmark = pytest.mark.gpu


def _devices(n_members):
    n = torch.cuda.device_count()
    return [i % n for i in range(n_members)]


def _frames_on(devs, arrays, w, h, fmt, memory):
    bufs, descs = [], []
    for i, a in enumerate(arrays):
        if memory == "device":
            t = torch.from_numpy(a.copy()).to(f"cuda:{devs[i % len(devs)]}")
        elif memory == "pinned":
            t = torch.from_numpy(a.copy()).pin_memory()
        else:
            t = a.copy()
        bufs.append(t)
        descs.append(frame_of(t, w, h, fmt))
    return bufs, descs


def _np(buf):
    return buf if isinstance(buf, np.ndarray) else buf.cpu().numpy()


@pytest.mark.parametrize("members", [1, 2, 3])
@pytest.mark.parametrize("memory", ["device", "host", "pinned"])
def test_group_round_robin_matches_oracle(orc, members, memory):
    devs = _devices(members)
    w, h, n = 320, 40, 7   # 7 frames over 2 or 3 members: uneven shares
    text = frames.cube_text_3d(9)
    lut = orc.Lut(text=text)
    srcs = [frames.frame_rand(w, h, 4, 100 + i).reshape(-1) for i in range(n)]
    with g.Group(devs) as grp:
        assert len(grp) == members
        grp.set_lut_from_cube(g.parse_cube(text))
        # colorlut, out of place
        ins, fin = _frames_on(devs, srcs, w, h, "RGBA", memory)
        outs, fout = _frames_on(devs, [np.zeros_like(s) for s in srcs], w, h, "RGBA", memory)
        grp.colorlut_batch(fin, fout)
        grp.synchronize()
        for i in range(n):
            assert np.array_equal(_np(outs[i]), orc.colorlut(lut, srcs[i], w, h)), ("colorlut", i)
        # hsvfilter, in place
        bufs, fr = _frames_on(devs, srcs, w, h, "BGRA", memory)
        grp.hsvfilter_batch(fr, g.HsvFilterParams(*util.CFG2))
        grp.synchronize()
        for i in range(n):
            assert np.array_equal(_np(bufs[i]), orc.hsvfilter(srcs[i].copy(), w, h, "BGRA", util.CFG2)), i
        # hsvdetector
        ins, fin = _frames_on(devs, srcs, w, h, "BGRx", memory)
        outs, fout = _frames_on(devs, [np.zeros_like(s) for s in srcs], w, h, "ARGB", memory)
        grp.hsvdetector_batch(fin, fout, g.HsvDetectorParams(*util.DET_CFG4))
        grp.synchronize()
        for i in range(n):
            assert np.array_equal(_np(outs[i]), orc.hsvdetector(srcs[i], w, h, "BGRx", "ARGB", util.DET_CFG4)), i
        # chain
        ins, fin = _frames_on(devs, srcs, w, h, "RGBA", memory)
        outs, fout = _frames_on(devs, [np.zeros_like(s) for s in srcs], w, h, "RGBA", memory)
        grp.chain_lut_hsv_batch(fin, fout, g.HsvFilterParams(*util.CFG2))
        grp.synchronize()
        for i in range(n):
            want = orc.hsvfilter(orc.colorlut(lut, srcs[i], w, h), w, h, "RGBA", util.CFG2)
            assert np.array_equal(_np(outs[i]), want), ("chain", i)
        # every member worked on its share only
        shares = [grp.member(m).stats()["frames"] for m in range(members)]
        want_share = [4 * len(range(m, n, members)) for m in range(members)]
        assert shares == want_share


def test_group_errors_and_options(orc):
    with pytest.raises(B200VFError):
        g.Group([])
    with pytest.raises(B200VFError):
        g.Group([torch.cuda.device_count() + 5])
    with g.Group(_devices(2)) as grp:
        w, h = 64, 8
        src = frames.frame_rand(w, h, 4, 1).reshape(-1)
        bufs, fr = _frames_on([0], [src, src], w, h, "RGBA", "host")
        outs, fo = _frames_on([0], [src, src], w, h, "RGBA", "host")
        with pytest.raises(B200VFError) as e:   # no LUT yet: the member's error comes through
            grp.colorlut_batch(fr, fo)
        assert "No LUT configured" in str(e.value) and "member 0" in str(e.value)
        grp.set_option("hsv.path", 1)
        assert [grp.member(m).get_option("hsv.path") for m in range(2)] == [1, 1]
        with pytest.raises(B200VFError):
            grp.set_option("no.such.option", 1)
        grp.colorlut_batch([], [])   # empty batch is fine


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_group_rejects_frames_on_the_wrong_device(orc):
    with g.Group([0, 1]) as grp:
        w, h = 64, 8
        src = frames.frame_rand(w, h, 4, 1).reshape(-1)
        # both frames on device 0: frame 1 belongs to member 1 (device 1)
        bufs, fr = _frames_on([0], [src, src], w, h, "RGBA", "device")
        with pytest.raises(B200VFError) as e:
            grp.hsvfilter_batch(fr, g.HsvFilterParams(*util.CFG2))
        assert "member 1" in str(e.value)


def test_group_calls_in_flight(orc):
    """\"host.async\" through the group: calls on pinned frames return once every member has queued
    its share; per-member tickets or group.synchronize() complete them."""
    devs = _devices(2)
    w, h, n = 640, 120, 6
    text = frames.cube_text_3d(9)
    lut = orc.Lut(text=text)
    with g.Group(devs) as grp:
        grp.set_lut_from_cube(g.parse_cube(text))
        grp.set_option("host.async", 1)
        batches = []
        for k in range(5):
            srcs = [frames.frame_rand(w, h, 4, 700 + 10 * k + i).reshape(-1) for i in range(n)]
            ins, fin = _frames_on(devs, srcs, w, h, "RGBA", "pinned")
            outs, fout = _frames_on(devs, [np.zeros_like(s) for s in srcs], w, h, "RGBA", "pinned")
            grp.colorlut_batch(fin, fout)
            tickets = [grp.member(m).host_ticket() for m in range(len(devs))]
            batches.append((srcs, ins, outs, tickets))
            if k:   # one call of latency: complete the previous batch through the members' tickets
                psrcs, _, pouts, ptickets = batches[k - 1]
                for m, t in enumerate(ptickets):
                    grp.member(m).host_wait(t)
                for i in range(n):
                    assert np.array_equal(_np(pouts[i]), orc.colorlut(lut, psrcs[i], w, h)), (k - 1, i)
        grp.synchronize()
        srcs, _, outs, _ = batches[-1]
        for i in range(n):
            assert np.array_equal(_np(outs[i]), orc.colorlut(lut, srcs[i], w, h)), i
        grp.set_option("host.async", 0)
