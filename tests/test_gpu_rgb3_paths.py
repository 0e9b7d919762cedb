"""3-byte pixel formats (RGB / BGR) through the function tables, end to end through the C ABI from
plain C (tests/cpp/rgb3_paths.c): hsvfilter in place, hsvdetector 3 -> 4 bytes and colorlut_convert
from 3-byte input, on geometries that take vf_map_tile3_staged_kernel (width % 128 == 0, 16-byte
aligned rows: 4K, 1080p, padded strides, one segment, rows not a multiple of 32) and on ones that do
not (other widths, 4-byte aligned and odd strides) — table path and compute path against the oracle,
device and system memory, batches."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build(tmp_path):
    import gst_plugins_rs_b200 as g  # noqa: F401  (makes sure the library is built)
    import oracle
    oracle.build()
    exe = tmp_path / "rgb3_paths"
    lib, orc = os.path.join(ROOT, "gst-plugins-rs_b200"), os.path.join(ROOT, "oracle")
    subprocess.run(["gcc", "-O2", "-std=c11", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
                    "-I" + orc, os.path.join(ROOT, "tests", "cpp", "rgb3_paths.c"), "-L" + lib, "-lb200vf",
                    "-L" + orc, "-loracle", "-lm", "-Wl,-rpath," + lib, "-Wl,-rpath," + orc, "-o", str(exe)],
                   check=True)
    return exe


def test_rgb3_harness_builds(tmp_path):
    """CPU: the harness compiles against the public header and the oracle's; without a device it
    says so and exits 3 (it never computes anything on the host instead)."""
    exe = build(tmp_path)
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 3 and "no device" in out.stdout, out.stdout + out.stderr


@pytest.mark.gpu
def test_rgb3_paths_equal_oracle(tmp_path):
    exe = build(tmp_path)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "ALL EQUAL" in out.stdout, out.stdout + out.stderr
