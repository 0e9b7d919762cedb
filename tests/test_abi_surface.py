"""The C-ABI library loads without a GPU and exports every symbol include/b200vf.h declares;
the Python prototype table matches the header; the product never touches oracle/."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200vf.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"B200VF_API\s+[^;(]*?\b(b200vf_\w+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    for must in ("b200vf_ctx_create", "b200vf_ctx_destroy", "b200vf_colorlut_set_lut",
                 "b200vf_colorlut_process", "b200vf_hsvfilter_process",
                 "b200vf_hsvdetector_process", "b200vf_cube_parse", "b200vf_last_error"):
        assert must in syms
    assert len(syms) >= 30


def test_library_exports_every_declared_symbol(vf):
    lib = vf._lib.load()
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in b200vf.h but not exported"


def test_python_prototypes_match_header(vf):
    assert sorted(vf._lib.PROTOTYPES) == declared_symbols()


def test_struct_layouts_match_header(vf):
    """sizeof/offsets of the plain structs the ABI passes (no torch types in the signatures)."""
    F = vf._lib.Frame
    assert ctypes.sizeof(F) == 32
    assert (F.data.offset, F.stride.offset, F.width.offset, F.height.offset, F.format.offset,
            F.memory.offset) == (0, 8, 16, 20, 24, 28)
    assert ctypes.sizeof(vf._lib.HsvFilterParams) == 20
    assert ctypes.sizeof(vf._lib.HsvDetectorParams) == 24
    assert ctypes.sizeof(vf._lib.Stats) == 32


def test_format_table(vf):
    lib = vf._lib.load()
    from gst_plugins_rs_b200.api import BYTES_PER_PIXEL, FORMATS
    for name, idx in FORMATS.items():
        assert lib.b200vf_format_name(idx).decode() == name
        assert lib.b200vf_format_from_name(name.encode()) == idx
        assert lib.b200vf_format_bytes_per_pixel(idx) == BYTES_PER_PIXEL[name]
    assert lib.b200vf_format_from_name(b"I420") == -1
    assert lib.b200vf_format_bytes_per_pixel(99) == 0
    assert b"sm_100a" in lib.b200vf_version()


def test_no_cpu_fallback_without_device(vf):
    """Without a CUDA device the product fails loudly (no oracle / CPU path behind the ABI)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from gst_plugins_rs_b200.api import ERR_NO_DEVICE, B200VFError
    with pytest.raises(B200VFError) as e:
        vf.Context(0)
    assert e.value.status == ERR_NO_DEVICE


def test_product_does_not_reference_oracle():
    """Only tests/, smoke() and bench.py may touch oracle/."""
    pkg = os.path.join(ROOT, "gst-plugins-rs_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp", ".rs")) or f == "Makefile":
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, flags=re.M), f
                assert "liboracle" not in src and "vf_oracle" not in src and "orc_" not in src, f
    so = os.path.join(pkg, "libb200vf.so")
    out = subprocess.run(["nm", "-D", so], capture_output=True, text=True).stdout
    assert "orc_" not in out
    deps = subprocess.run(["ldd", so], capture_output=True, text=True).stdout
    assert "oracle" not in deps


def test_only_abi_symbols_are_exported():
    so = os.path.join(ROOT, "gst-plugins-rs_b200", "libb200vf.so")
    out = subprocess.run(["nm", "-D", "--defined-only", so], capture_output=True, text=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    extra = {s for s in exported if not s.startswith("b200vf_") and s not in ("_init", "_fini")}
    assert not extra, f"non-ABI symbols exported: {sorted(extra)[:5]}"
    assert set(declared_symbols()) <= exported


def test_header_compiles_as_c_and_links(tmp_path):
    """include/b200vf.h from plain C11 (-Wall -Wextra -Werror -pedantic), linked against the
    library: struct layouts, format table, the reference's parser KAT, no-device behaviour."""
    src = os.path.join(ROOT, "tests", "c_abi", "abi_c_consumer.c")
    exe = tmp_path / "abi_c_consumer"
    libdir = os.path.join(ROOT, "gst-plugins-rs_b200")
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-pedantic",
                    "-I", os.path.join(ROOT, "include"), "-o", str(exe), src, "-L", libdir, "-lb200vf",
                    f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "abi_c_consumer: ok" in out.stdout


def test_group_fails_loudly_without_device(vf):
    """The frame-parallel group is N contexts: without a CUDA device it cannot be created either."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from gst_plugins_rs_b200.api import ERR_INVALID_ARG, ERR_NO_DEVICE, B200VFError
    with pytest.raises(B200VFError) as e:
        vf.Group([0, 1])
    assert e.value.status == ERR_NO_DEVICE
    with pytest.raises(B200VFError) as e:
        vf.Group([])
    assert e.value.status == ERR_INVALID_ARG


def test_element_library_exports_its_c_header():
    """libb200vf_elements.so (the C++ element mirror) exports every function of elements/vf_elements_c.h
    and loads without a GPU."""
    from gst_plugins_rs_b200 import elements
    hdr = open(os.path.join(ROOT, "gst-plugins-rs_b200", "elements", "vf_elements_c.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"B200VF_API\s+[^;(]*?\b(b200vf_element_\w+)\s*\(", hdr)))
    assert len(names) >= 16 and "b200vf_element_generate_output" in names
    lib = elements.load()
    for n in names:
        assert hasattr(lib, n), f"{n} declared in vf_elements_c.h but not exported"


def test_queued_operation_protocol_without_a_gpu():
    """The queued pair's bookkeeping needs no device: nothing queued -> no output, latency bounds."""
    from gst_plugins_rs_b200 import elements
    e = elements.Element("hsvfilter")
    assert e.generate_output() is None and e.drain() == []
    e.set_frames_in_flight(3)
    with pytest.raises(elements.ElementError):
        e.set_frames_in_flight(15)
    e.set_frames_in_flight(0)


def test_integration_doc_has_a_row_for_every_entry_point():
    """INTEGRATION.md §1 says what each exported function replaces in the reference: none may be
    missing (names may be written with `*` wildcards or `a[_b]` options)."""
    import fnmatch
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    sec = doc[doc.index("## 1."):doc.index("## 2.")]
    tokens = set(re.findall(r"`([^`]+)`", sec))
    names = set()
    for t in tokens:
        t = t.split("(")[0].strip()
        if not (t.startswith("b200vf_") or t.startswith("*_")):
            continue
        m = re.match(r"^(.*)\[(\w+)\](.*)$", t)           # b200vf_cube_parse[_file]
        names.update([m.group(1) + m.group(3), m.group(1) + m.group(2) + m.group(3)] if m else [t])
    missing = [s for s in declared_symbols()
               if s not in names and not any(fnmatch.fnmatch(s, n) for n in names if "*" in n)]
    assert not missing, missing
