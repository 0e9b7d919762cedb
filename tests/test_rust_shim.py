"""The Rust shims under gst-plugins-rs_b200/rust cannot be compiled in this image (no rustc, no
GStreamer).  What can be checked without a compiler is checked: the FFI block declares exactly the
header's symbols, the #[repr(C)] structs have the header's fields in the header's order, and the
element surface written out in the imp.rs files equals the reference's
(tests/golden/element_surface.json, extracted from docs/plugins/gst_plugins_cache.json)."""
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUST = os.path.join(ROOT, "gst-plugins-rs_b200", "rust")
HEADER = os.path.join(ROOT, "include", "b200vf.h")


def _read(*parts):
    return open(os.path.join(RUST, *parts)).read()


def _header():
    return re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)


def test_workspace_is_self_contained():
    for rel in ("Cargo.toml", "b200vf-sys/Cargo.toml", "b200vf-sys/build.rs", "b200vf-sys/src/lib.rs",
                "colorlut/Cargo.toml", "colorlut/build.rs", "colorlut/src/lib.rs",
                "colorlut/src/colorlut/mod.rs", "colorlut/src/colorlut/imp.rs", "hsv/Cargo.toml",
                "hsv/build.rs", "hsv/src/lib.rs", "hsv/src/shared.rs", "hsv/src/hsvfilter/mod.rs",
                "hsv/src/hsvfilter/imp.rs", "hsv/src/hsvdetector/mod.rs", "hsv/src/hsvdetector/imp.rs"):
        assert os.path.exists(os.path.join(RUST, rel)), rel
    ws = _read("Cargo.toml")
    assert 'members = ["b200vf-sys", "colorlut", "hsv"]' in ws
    for crate, lib, plugin in (("colorlut", "gstcolorlut", "colorlut"), ("hsv", "gsthsv", "hsv")):
        assert f'name = "{lib}"' in _read(crate, "Cargo.toml")
        src = _read(crate, "src", "lib.rs")
        assert "gst::plugin_define!(" in src and re.search(r"plugin_define!\(\s*%s," % plugin, src)
    # every `mod x;` resolves to a file of the crate
    for crate in ("colorlut", "hsv"):
        for m in re.findall(r"^mod (\w+);", _read(crate, "src", "lib.rs"), flags=re.M):
            base = os.path.join(RUST, crate, "src", m)
            assert os.path.exists(base + ".rs") or os.path.exists(os.path.join(base, "mod.rs")), m


def test_ffi_block_declares_exactly_the_header_symbols():
    declared = sorted(set(re.findall(r"B200VF_API\s+[^;(]*?\b(b200vf_\w+)\s*\(", _header())))
    lib = _read("b200vf-sys", "src", "lib.rs")
    block = re.search(r'extern "C" \{(.*?)\n\}', lib, flags=re.S).group(1)
    rust = sorted(re.findall(r"pub fn (b200vf_\w+)\(", block))
    assert rust == declared
    assert len(rust) == len(set(rust))


def test_ffi_argument_counts_match_the_header():
    hdr = {}
    for ret, name, args in re.findall(r"B200VF_API\s+([^;(]*?)\b(b200vf_\w+)\s*\(([^;]*?)\)\s*;", _header()):
        args = " ".join(args.split())
        hdr[name] = 0 if args in ("", "void") else len(args.split(","))
    lib = _read("b200vf-sys", "src", "lib.rs")
    for name, args in re.findall(r"pub fn (b200vf_\w+)\(([^)]*)\)", lib):
        n = 0 if not args.strip() else len([a for a in args.split(",") if a.strip()])
        assert n == hdr[name], name


def test_repr_c_structs_follow_the_header():
    hdr = _header()
    lib = _read("b200vf-sys", "src", "lib.rs")
    for name in ("b200vf_frame", "b200vf_stats", "b200vf_cube", "b200vf_hsvfilter_params",
                 "b200vf_hsvdetector_params", "b200vf_pool_config", "b200vf_pool_stats"):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), hdr, flags=re.S).group(1)
        c_fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            names = decl.split(",")
            first = re.match(r".*?(\w+)(\[\d+\])?$", names[0].strip()).group(1)
            c_fields += [first] + [re.match(r"\*?\s*(\w+)", n.strip()).group(1) for n in names[1:]]
        m = re.search(r"#\[repr\(C\)\][^{]*?pub struct %s \{(.*?)\n\}" % name, lib, flags=re.S)
        rust_fields = re.findall(r"pub (\w+):", m.group(1))
        assert rust_fields == c_fields, name


def _surface_of(src):
    """Properties and format lists as written in an imp.rs of the shims."""
    props = {}
    for m in re.finditer(r'glib::ParamSpecString::builder\("([\w-]+)"\)', src):
        props[m.group(1)] = {"type": "gchararray"}
    return props


def test_element_surface_matches_the_reference():
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "element_surface.json")))
    lut = _read("colorlut", "src", "colorlut", "imp.rs")
    flt = _read("hsv", "src", "hsvfilter", "imp.rs")
    det = _read("hsv", "src", "hsvdetector", "imp.rs")
    for name, src, mod in (("colorlut", lut, _read("colorlut", "src", "colorlut", "mod.rs")),
                           ("hsvfilter", flt, _read("hsv", "src", "hsvfilter", "mod.rs")),
                           ("hsvdetector", det, _read("hsv", "src", "hsvdetector", "mod.rs"))):
        g = gold[name]
        assert 'const NAME: &\'static str = "%s";' % g["gtype"] in src
        assert "type ParentType = gst_video::VideoFilter;" in src
        assert '"%s",' % g["klass"] in src
        assert re.search(r'Element::register\(Some\(plugin\), "%s", gst::Rank::NONE' % name, mod)
    # colorlut: `location` string, mutable in READY; formats in the reference's order
    assert re.search(r'ParamSpecString::builder\("location"\)[^;]*?\.mutable_ready\(\)', lut, flags=re.S)
    assert "VideoFormat::Rgba64Le, VideoFormat::Rgba64Be, VideoFormat::Rgba" in lut
    assert "BaseTransformMode::NeverInPlace" in lut and "BaseTransformMode::NeverInPlace" in det
    assert "BaseTransformMode::AlwaysInPlace" in flt
    fmt = {"RGBx": "Rgbx", "xRGB": "Xrgb", "BGRx": "Bgrx", "xBGR": "Xbgr", "RGBA": "Rgba", "ARGB": "Argb",
           "BGRA": "Bgra", "ABGR": "Abgr", "RGB": "Rgb", "BGR": "Bgr"}

    def fmt_list(src, const):
        body = re.search(r"const %s: \[VideoFormat; \d+\] =?\s*\[(.*?)\];" % const, src, flags=re.S).group(1)
        return re.findall(r"VideoFormat::(\w+)", body)
    assert fmt_list(flt, "FORMATS") == [fmt[f] for f in gold["hsvfilter"]["sink_formats"]]
    assert fmt_list(det, "INPUT_FORMATS") == [fmt[f] for f in gold["hsvdetector"]["sink_formats"]]
    assert fmt_list(det, "OUTPUT_FORMATS") == [fmt[f] for f in gold["hsvdetector"]["src_formats"]]
    # float properties: name, default, range — the tables in the two imp.rs files
    for name, src in (("hsvfilter", flt), ("hsvdetector", det)):
        table = re.search(r"const FLOAT_PROPS: .*? = \[(.*?)\n\];", src, flags=re.S).group(1)
        rows = re.findall(r'\(\s*"([\w-]+)",\s*"[^"]*",\s*"[^"]*",\s*([-\d.]+),?\s*(None|Some\(\(([-\d.]+), ([-\d.]+)\)\))?,?\s*\)',
                          table, flags=re.S)
        assert [r[0] for r in rows] == list(gold[name]["properties"]), name
        for pname, default, rng, lo, hi in rows:
            g = gold[name]["properties"][pname]
            assert g["type"] == "gfloat" and g["mutable"] == "playing"
            assert float(default) == float(g["default"]), pname
            if lo:
                assert (float(lo), float(hi)) == (float(g["min"]), float(g["max"])), pname
            else:  # unbounded in the shim = the whole f32 range in the reference
                assert float(g["max"]) > 3.4e38 and float(g["min"]) < -3.4e38, pname
        assert ".mutable_playing()" in src
    # the one added property defaults to the reference's behaviour (device 0)
    for src in (lut, _read("hsv", "src", "shared.rs")):
        assert re.search(r'ParamSpecInt::builder\("device"\).*?\.default_value\(0\)', src, flags=re.S)
