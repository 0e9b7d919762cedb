"""Element-layer conformance: the C++ mirror of the reference elements must expose exactly the
surface recorded in docs/plugins/gst_plugins_cache.json (committed extract:
tests/golden/element_surface.json), GObject property semantics, start/stop errors
(colorlut/imp.rs:168-199) and hsvdetector's transform_caps (hsvdetector/imp.rs:386-419)."""
import json
import os

import numpy as np
import pytest

import util
from gst_plugins_rs_b200 import elements, frames
from gst_plugins_rs_b200.api import frame_of

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "element_surface.json")))


@pytest.mark.parametrize("name", ["colorlut", "hsvfilter", "hsvdetector"])
def test_surface_matches_docs_cache(name):
    got, want = elements.describe(name), GOLDEN[name]
    for key in ("gtype", "parent", "klass", "rank", "plugin", "filename", "license"):
        assert got[key] == want[key], key
    assert got["sink_formats"] == want["sink_formats"]   # order matters for default fixation
    assert got["src_formats"] == want["src_formats"]
    assert sorted(got["properties"]) == sorted(want["properties"])
    for pname, spec in want["properties"].items():
        for k, v in spec.items():
            assert got["properties"][pname][k] == v, (pname, k)


def test_base_transform_modes():
    assert elements.describe("colorlut")["mode"] == "NeverInPlace"      # colorlut/imp.rs:163
    assert elements.describe("hsvfilter")["mode"] == "AlwaysInPlace"    # hsvfilter/imp.rs:316
    assert elements.describe("hsvdetector")["mode"] == "NeverInPlace"   # hsvdetector/imp.rs:381


def test_unknown_factory():
    with pytest.raises(ValueError):
        elements.Element("d3d12colorlut")


def test_property_defaults_and_validation():
    e = elements.Element("hsvfilter")
    assert [e.get_property(p) for p in ("hue-shift", "saturation-mul", "saturation-off",
                                        "value-mul", "value-off")] == [0.0, 1.0, 0.0, 1.0, 0.0]
    assert e.set_property("hue-shift", -1e38) and e.get_property("hue-shift") == np.float32(-1e38)
    assert not e.set_property("no-such-prop", 1.0)
    d = elements.Element("hsvdetector")
    want = {"hue-ref": 0.0, "hue-var": 10.0, "saturation-ref": 0.0, "saturation-var": 0.15,
            "value-ref": 0.0, "value-var": 0.3}
    for k, v in want.items():
        assert d.get_property(k) == np.float32(v)
    # ranged ParamSpecFloat: out-of-range is refused and the old value kept
    assert not d.set_property("hue-var", 181.0) and d.get_property("hue-var") == 10.0
    assert not d.set_property("saturation-ref", -0.1)
    assert d.set_property("hue-var", 180.0) and d.get_property("hue-var") == 180.0
    assert d.set_property("hue-ref", 1e30)     # unbounded
    c = elements.Element("colorlut")
    assert c.get_property("location") is None  # default NULL
    assert c.set_property("location", "/tmp/x.cube") and c.get_property("location") == "/tmp/x.cube"
    assert c.set_property("location", None) and c.get_property("location") is None


def test_hsvdetector_transform_caps():
    d = elements.Element("hsvdetector")
    ins = ["RGBx", "xRGB", "BGRx", "xBGR", "RGB", "BGR"]
    outs = ["RGBA", "ARGB", "BGRA", "ABGR"]
    assert d.transform_caps("sink", ["BGRx"]) == outs        # sink → src: all 4, RGBA first
    assert d.transform_caps("src", ["ARGB"]) == ins          # src → sink: all 6
    # filter caps, First mode: the filter's order wins
    assert d.transform_caps("sink", ["BGRx"], ["ABGR", "I420", "RGBA"]) == ["ABGR", "RGBA"]
    assert d.transform_caps("src", ["RGBA"], ["BGR", "RGBA"]) == ["BGR"]
    # colorlut / hsvfilter: default GstVideoFilter behaviour (same caps both sides)
    f = elements.Element("hsvfilter")
    assert f.transform_caps("sink", ["RGB", "BGRA"]) == ["RGB", "BGRA"]


def test_colorlut_start_errors_without_gpu(tmp_path):
    """start(): missing location → ResourceError::Settings (imp.rs:175-180) — checked before any
    device work, so this also runs on the CPU-only CI."""
    c = elements.Element("colorlut")
    with pytest.raises(elements.ElementError) as e:
        c.start()
    assert e.value.domain == "Settings" and e.value.message == "LUT file location is not configured"
    # transform_frame before a successful start: FlowError::Error + "No LUT configured"
    buf = np.zeros(16, np.uint8)
    f = frame_of(buf, 4, 1, "RGBA")
    assert c.transform_frame(f, f) == elements.FLOW_ERROR and "No LUT configured" in c.message()


@pytest.mark.gpu
def test_colorlut_element_pipeline(orc, tmp_path):
    """videotestsrc-like frames ! colorlut location=… ! sink, host memory (cfg1 shape)."""
    import torch
    assert torch.cuda.is_available()
    c = elements.Element("colorlut")
    c.set_property("location", tmp_path / "missing.cube")
    with pytest.raises(elements.ElementError) as e:
        c.start()
    assert e.value.domain == "Read" and "Failed to parse LUT file" in e.value.message
    path = tmp_path / "lut33.cube"
    text = frames.cube_text_3d(33)
    path.write_text(text)
    c.set_property("location", path)
    c.start()
    lut = orc.Lut(text=text)
    w, h = 1920, 1080
    for i in range(3):
        src = frames.frame_bars(w, h).reshape(-1).copy()
        src[::7] ^= i  # vary the frames a little
        dst = np.zeros_like(src)
        assert c.transform_frame(frame_of(src, w, h, "RGBA"), frame_of(dst, w, h, "RGBA")) == 0
        assert np.array_equal(dst, orc.colorlut(lut, src, w, h))
    # wrong format → FlowError (the reference would never negotiate it)
    assert c.transform_frame(frame_of(src, w, h, "BGRA"), frame_of(dst, w, h, "BGRA")) == elements.FLOW_ERROR
    c.stop()
    assert c.transform_frame(frame_of(src, w, h, "RGBA"), frame_of(dst, w, h, "RGBA")) == elements.FLOW_ERROR


@pytest.mark.gpu
def test_hsv_elements_pipeline_and_live_property_change(orc):
    """hsvfilter: a property changed between frames applies to the next frame (mutable-playing,
    snapshot per frame, hsvfilter/imp.rs:85); hsvdetector with cfg4 settings."""
    w, h = 1280, 720
    f = elements.Element("hsvfilter")
    src = frames.frame_rand(w, h, 4, 9).reshape(-1)
    buf = src.copy()
    assert f.transform_frame_ip(frame_of(buf, w, h, "BGRA")) == 0
    assert np.array_equal(buf, orc.hsvfilter(src, w, h, "BGRA", util.IDENTITY))
    for name, v in zip(("hue-shift", "saturation-mul", "saturation-off", "value-mul", "value-off"),
                       util.CFG2):
        assert f.set_property(name, v)
    buf = src.copy()
    assert f.transform_frame_ip(frame_of(buf, w, h, "BGRA")) == 0
    assert np.array_equal(buf, orc.hsvfilter(src, w, h, "BGRA", util.CFG2))
    d = elements.Element("hsvdetector", hue_ref=120.0, hue_var=30.0, saturation_ref=0.6,
                         saturation_var=0.4, value_ref=0.6, value_var=0.4)
    out = np.zeros(w * h * 4, np.uint8)
    assert d.transform_frame(frame_of(src, w, h, "BGRx"), frame_of(out, w, h, "RGBA")) == 0
    assert np.array_equal(out, orc.hsvdetector(src, w, h, "BGRx", "RGBA", util.DET_CFG4))
    # a format pair outside the caps is refused
    assert d.transform_frame(frame_of(src, w, h, "RGBA"), frame_of(out, w, h, "RGBA")) == elements.FLOW_ERROR


@pytest.mark.gpu
def test_cfg1_pipeline_example_cpp(tmp_path):
    """BASELINE configs[0] as a C++ program over the element layer (no Python in the data path):
    videotestsrc-like buffers ! colorlut(33^3) 1920x1080 RGBA ! fakesink."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "examples", "cfg1_pipeline")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", os.path.join(root, "examples")], check=True)
    cube = tmp_path / "lut33.cube"
    cube.write_text(frames.cube_text_3d(33))
    out = subprocess.run([exe, str(cube), "60"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    res = json.loads(out.stdout.strip().splitlines()[-1])
    assert res["frames_per_s"] > 0 and "colorlut" in res["pipeline"]
    # the same pipeline with the element's proposed (page-locked) pools on both pads
    pooled = subprocess.run([exe, str(cube), "60", "1920", "1080", "4", "8388608", "pool"],
                            capture_output=True, text=True, timeout=300)
    assert pooled.returncode == 0, pooled.stderr
    res_pool = json.loads(pooled.stdout.strip().splitlines()[-1])
    assert "page-locked pool" in res_pool["memory"] and res_pool["checksum"] == json.loads(
        subprocess.run([exe, str(cube), "60"], capture_output=True, text=True).stdout.strip()
        .splitlines()[-1])["checksum"]
    # missing LUT file → start() fails like the reference element (ResourceError::Read)
    bad = subprocess.run([exe, str(tmp_path / "nope.cube"), "1"], capture_output=True, text=True)
    assert bad.returncode == 4 and "Failed to parse LUT file" in bad.stderr


@pytest.mark.gpu
def test_queued_operation_holds_one_frame_back(orc, tmp_path):
    """submit_input_frame / generate_output (BaseTransform's queued pair): with one frame in
    flight the element hands out buffer k-1 — complete, the oracle's pixels — when buffer k is
    submitted; drain hands out the rest; pageable frames go through the same calls."""
    import torch
    w, h = 1280, 720
    text = frames.cube_text_3d(17)
    path = tmp_path / "lut17.cube"
    path.write_text(text)
    lut = orc.Lut(text=text)
    c = elements.Element("colorlut", location=path)
    c.set_frames_in_flight(1)          # before start: applied when the context exists
    c.start()
    n = 6
    srcs = [frames.frame_rand(w, h, 4, 50 + i).reshape(-1) for i in range(n)]
    h_in = [torch.from_numpy(s.copy()).pin_memory() for s in srcs]
    h_out = [torch.zeros(w * h * 4, dtype=torch.uint8).pin_memory() for _ in srcs]
    by_ptr = {t.data_ptr(): i for i, t in enumerate(h_out)}
    order = []

    def sink(done):
        i = by_ptr[done.data]
        assert np.array_equal(h_out[i].numpy(), orc.colorlut(lut, srcs[i], w, h)), i
        order.append(i)

    for i in range(n):
        assert c.submit_input_frame(frame_of(h_in[i], w, h, "RGBA"), frame_of(h_out[i], w, h, "RGBA")) == 0
        done = c.generate_output()
        assert (done is None) == (i == 0)
        if done is not None:
            sink(done)
    for done in c.drain():
        sink(done)
    assert order == list(range(n)) and c.drain() == []
    with pytest.raises(elements.ElementError):
        c.set_frames_in_flight(15)
    c.stop()
    # stopped with frames still held back (a device-following restart in mid-stream): nothing is
    # lost, the frames are complete and come out of drain()
    c.set_frames_in_flight(2)
    c.start()
    order.clear()
    for i in range(2):
        h_out[i].zero_()
        assert c.submit_input_frame(frame_of(h_in[i], w, h, "RGBA"), frame_of(h_out[i], w, h, "RGBA")) == 0
        assert c.generate_output() is None
    c.stop()
    for done in c.drain():
        sink(done)
    assert order == [0, 1]

    # in place, pageable, two frames held back: the calls are synchronous underneath, same protocol
    f = elements.Element("hsvfilter")
    f.set_frames_in_flight(2)
    for name, v in zip(("hue-shift", "saturation-mul", "saturation-off", "value-mul", "value-off"), util.CFG2):
        assert f.set_property(name, v)
    bufs = [s.copy() for s in srcs[:4]]
    got = []
    for b in bufs:
        assert f.submit_input_frame(frame_of(b, w, h, "RGBA")) == 0
        done = f.generate_output()
        if done is not None:
            got.append(done.data)
    got += [d.data for d in f.drain()]
    assert got == [b.ctypes.data for b in bufs]
    for b, s in zip(bufs, srcs):
        assert np.array_equal(b, orc.hsvfilter(s, w, h, "RGBA", util.CFG2))
    # back to the reference's synchronous behaviour
    f.set_frames_in_flight(0)
    b = srcs[0].copy()
    assert f.submit_input_frame(frame_of(b, w, h, "RGBA")) == 0
    assert f.generate_output().data == b.ctypes.data
