"""device frame pool (b200vf_pool_*), pointer classification and the
CUDA-memory element variants' negotiation, modelled on d3d12colorlut
(video/colorlut/src/d3d12colorlut/imp.rs:236-266, 349-542).  The pixel results of the
device-resident pipeline are compared with the oracle."""
import ctypes as C
import json
import os
import subprocess
import threading
import time

import numpy as np
import pytest

import gst_plugins_rs_b200 as g
from gst_plugins_rs_b200 import _lib, api, elements, frames

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "gst-plugins-rs_b200")


def _build_check(tmp_path):
    exe = tmp_path / "negotiation_check"
    subprocess.run(["g++", "-O1", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-o", str(exe),
                    os.path.join(ROOT, "tests", "cpp", "negotiation_check.cpp"), "-L", LIBDIR,
                    "-lb200vf_elements", "-lb200vf", f"-Wl,-rpath,{LIBDIR}"], check=True)
    return str(exe)


def test_negotiation_logic_without_device(tmp_path):
    out = subprocess.run([_build_check(tmp_path), "cpu"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "negotiation_check cpu: ok" in out.stdout


def test_cuda_variants_surface():
    for base in ("colorlut", "hsvfilter", "hsvdetector"):
        b, c = elements.describe(base), elements.describe("cuda" + base)
        assert c["plugin"] == "b200vf" and c["rank"] == "none" and c["mode"] == b["mode"]
        assert c["sink_formats"] == b["sink_formats"] and c["src_formats"] == b["src_formats"]
        assert c["sink_features"] == c["src_features"] == ["memory:CUDAMemory"]
        assert "sink_features" not in b          # the reference elements stay system-memory
        assert c["properties"] == b["properties"]
        assert c["gtype"] == {"colorlut": "GstCudaColorLut", "hsvfilter": "GstCudaHsvFilter",
                              "hsvdetector": "GstCudaHsvDetector"}[base]


def test_pool_argument_validation():
    lib = _lib.load()
    pool = C.c_void_p()
    ok_cfg = _lib.PoolConfig(64, 32, 0, 0, 0)
    assert lib.b200vf_pool_create(0, None, C.byref(pool)) == api.ERR_INVALID_ARG
    assert lib.b200vf_pool_create(0, C.byref(ok_cfg), None) == api.ERR_INVALID_ARG
    for cfg, want in ((_lib.PoolConfig(64, 32, 99, 0, 0), api.ERR_UNSUPPORTED_FORMAT),
                      (_lib.PoolConfig(0, 32, 0, 0, 0), api.ERR_INVALID_ARG),
                      (_lib.PoolConfig(64, 0, 0, 0, 0), api.ERR_INVALID_ARG),
                      (_lib.PoolConfig(64, 32, 0, 4, 2), api.ERR_INVALID_ARG)):
        assert lib.b200vf_pool_create(0, C.byref(cfg), C.byref(pool)) == want
        assert not pool.value
    assert lib.b200vf_pool_create(10 ** 6, C.byref(ok_cfg), C.byref(pool)) == api.ERR_NO_DEVICE
    assert b"pool_create" in lib.b200vf_last_error(None) or b"cudaGetDeviceCount" in lib.b200vf_last_error(None)
    lib.b200vf_pool_destroy(None)  # no-op
    f = _lib.Frame()
    assert lib.b200vf_pool_acquire(None, 0, C.byref(f)) == api.ERR_INVALID_ARG
    assert lib.b200vf_pool_release(None, C.byref(f), None) == api.ERR_INVALID_ARG
    assert lib.b200vf_pool_device(None) == -1
    mem, dev = C.c_uint32(), C.c_int()
    assert lib.b200vf_pointer_info(None, C.byref(mem), C.byref(dev)) == api.ERR_INVALID_ARG


# --------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_pool_geometry_recycling_and_limits():
    with api.DevicePool(0, 1920, 1080, "RGBA", min_buffers=2, max_buffers=3) as pool:
        st = pool.stats()
        assert st == {"allocated": 2, "outstanding": 0, "frame_bytes": 1920 * 4 * 1080,
                      "stride": 1920 * 4}
        a, b, c = pool.acquire(), pool.acquire(), pool.acquire()
        assert len({a.data, b.data, c.data}) == 3 and pool.stats()["allocated"] == 3
        for f in (a, b, c):
            assert (f.width, f.height, f.format, f.memory, f.stride) == (1920, 1080, 0, api.MEM_DEVICE, 7680)
            assert f.data % 256 == 0
            assert api.pointer_info(f.data) == (api.MEM_DEVICE, 0)
        with pytest.raises(api.B200VFError) as ei:   # GST_BUFFER_POOL_ACQUIRE_FLAG_DONTWAIT at max
            pool.acquire(dont_wait=True)
        assert ei.value.status == api.ERR_NOMEM
        pool.release(b)
        assert pool.acquire().data == b.data         # recycled, nothing new allocated
        assert pool.stats() == {**st, "allocated": 3, "outstanding": 3}
        # a blocking acquire returns as soon as another thread releases a frame
        got = {}
        t = threading.Thread(target=lambda: got.setdefault("f", pool.acquire()))
        t.start()
        time.sleep(0.2)
        assert t.is_alive()
        pool.release(c)
        t.join(10)
        assert not t.is_alive() and got["f"].data == c.data
        foreign = _lib.Frame(a.data + 16, a.stride, a.width, a.height, a.format, a.memory)
        with pytest.raises(api.B200VFError) as ei:
            pool.release(foreign)
        assert ei.value.status == api.ERR_INVALID_ARG
    # rows that are not a multiple of 16 bytes are padded to 256 (row starts stay vector-aligned)
    with api.DevicePool(0, 1001, 7, "RGB") as pool:
        assert pool.stats()["stride"] == 3072 and pool.stats()["allocated"] == 0
        assert pool.acquire().stride == 3072
    with api.DevicePool(0, 1001, 7, "RGBA64_LE") as pool:
        assert pool.stats()["stride"] == 8192        # 8008 is not a multiple of 16
    with api.DevicePool(0, 1002, 7, "RGBA64_LE") as pool:
        assert pool.stats()["stride"] == 8016        # tight: contiguous frames, long-row path
    assert api.pointer_info(np.zeros(16, np.uint8).ctypes.data) == (api.MEM_HOST, -1)


@pytest.mark.gpu
def test_pool_frames_through_the_kernels_match_oracle(orc):
    """hsvdetector from one pool into another (padded stride), released with the context
    stream: the recycled frame is only handed out after the kernel has finished."""
    w, h = 1001, 37
    src = frames.frame_rand(w, h, 4, 7)
    want = orc.hsvdetector(src, w, h, "BGRx", "RGBA", (120.0, 30.0, 0.6, 0.4, 0.6, 0.4))
    with g.Context() as ctx, api.DevicePool(0, w, h, "BGRx") as pin, \
            api.DevicePool(0, w, h, "RGBA", max_buffers=1) as pout:
        params = g.HsvDetectorParams(120.0, 30.0, 0.6, 0.4, 0.6, 0.4)
        for _ in range(3):
            fi, fo = pin.acquire(), pout.acquire()
            assert fi.stride == 4096 and fo.stride == 4096
            host = np.zeros((h, fi.stride), np.uint8)
            host[:, :w * 4] = src.reshape(h, w * 4)
            ctx._check(ctx.lib.b200vf_memcpy(ctx.h, fi.data, host.ctypes.data, host.size, 0))
            ctx.hsvdetector(fi, fo, params)
            pin.release(fi, ctx.get_stream())
            out = np.zeros((h, fo.stride), np.uint8)
            ctx._check(ctx.lib.b200vf_memcpy(ctx.h, out.ctypes.data, fo.data, out.size, 1))
            pout.release(fo, ctx.get_stream())
            assert np.array_equal(out[:, :w * 4].reshape(-1), want)
        assert pin.stats()["allocated"] == 1 and pout.stats()["allocated"] == 1


@pytest.mark.gpu
def test_pinned_host_pool_frames_are_system_memory(orc):
    """host_pinned pools hand out page-locked SYSTEM memory: plain pointers the CPU writes, which the
    host path copies without a bounce (b200vf_pool_config.host_pinned)."""
    w, h = 641, 48
    src = frames.frame_rand(w, h, 4, 9)
    want = np.asarray(orc.hsvfilter(src.copy(), w, h, "BGRA", (200.0, 1.5, -0.2, 0.7, 0.1))).reshape(-1)
    with g.Context() as ctx, api.DevicePool(0, w, h, "BGRA", host_pinned=True) as pool:
        f = pool.acquire()
        assert f.memory == api.MEM_HOST and f.stride == 2816          # 2564 padded to 256
        assert pool.stats()["frame_bytes"] == f.stride * h
        assert api.pointer_info(f.data) == (api.MEM_HOST, -1)
        view = np.ctypeslib.as_array(C.cast(f.data, C.POINTER(C.c_uint8)), shape=(h, f.stride))
        for _ in range(2):
            view[:, :w * 4] = src.reshape(h, w * 4)
            view[:, w * 4:] = 0x5A                                      # padding must stay untouched
            before = ctx.stats()["h2d_bytes"]
            ctx.hsvfilter(f, g.HsvFilterParams(200.0, 1.5, -0.2, 0.7, 0.1))
            assert ctx.stats()["h2d_bytes"] - before == w * 4 * h
            assert np.array_equal(view[:, :w * 4].reshape(-1), want)
            assert (view[:, w * 4:] == 0x5A).all()
        pool.release(f)
        assert pool.acquire().data == f.data


@pytest.mark.gpu
def test_negotiation_on_device(tmp_path):
    cube = tmp_path / "lut17.cube"
    cube.write_text(frames.cube_text_3d(17))
    out = subprocess.run([_build_check(tmp_path), "gpu", str(cube)], capture_output=True, text=True,
                         timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "negotiation_check gpu: ok" in out.stdout


def _run_pipeline(tmp_path, orc, w, h, n, hue, extra=()):
    exe = os.path.join(ROOT, "examples", "cuda_memory_pipeline")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", os.path.join(ROOT, "examples")], check=True)
    cube_text = frames.cube_text_3d(17)
    cube = tmp_path / "lut17.cube"
    cube.write_text(cube_text)
    src = np.concatenate([frames.frame_rand(w, h, 4, 100 + i).reshape(-1) for i in range(n)])
    (tmp_path / "in.raw").write_bytes(src.tobytes())
    res = subprocess.run([exe, str(cube), str(tmp_path / "in.raw"), str(tmp_path / "out.raw"),
                          str(w), str(h), str(n), str(hue), *map(str, extra)],
                         capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr
    got = np.fromfile(tmp_path / "out.raw", np.uint8)
    lut = orc.Lut(text=cube_text)
    per = w * h * 4
    for i in range(n):
        mid = orc.colorlut(lut, src[i * per:(i + 1) * per], w, h, "RGBA")
        want = orc.hsvfilter(mid, w, h, "RGBA", (hue, 1.0, 0.0, 1.0, 0.0))
        assert np.array_equal(got[i * per:(i + 1) * per], want), f"frame {i}"
    return json.loads(res.stdout.strip().splitlines()[-1])


@pytest.mark.gpu
def test_cuda_memory_pipeline_example_matches_oracle(tmp_path, orc):
    """upload ! cudacolorlut ! cudahsvfilter ! download with negotiated pools, frames resident in
    HBM between the elements: bit-exact with colorlut → hsvfilter of the oracle, and the pools
    recycle (one frame each for a synchronous sink)."""
    st = _run_pipeline(tmp_path, orc, 640, 360, 6, 40.0)
    assert st["frames"] == 6 and st["device"] == 0 and st["downstream_negotiations"] == 1
    assert st["in_pool_outstanding"] == 0 and st["out_pool_outstanding"] == 0
    assert 1 <= st["in_pool_allocated"] <= 2 and st["out_pool_allocated"] == 1


@pytest.mark.gpu
def test_elements_follow_the_device_of_incoming_memory(tmp_path, orc):
    """d3d12colorlut/imp.rs:494-542: buffers arriving from another device move the element
    there (context + LUT recreated, downstream allocation renegotiated).  Needs 2 GPUs."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    st = _run_pipeline(tmp_path, orc, 320, 200, 4, -75.0, extra=(1,))
    assert st["device"] == 1 and st["downstream_negotiations"] == 2
