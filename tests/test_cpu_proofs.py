"""Exhaustive CPU proofs of the division shortcuts in csrc/vf_math.cuh (see
tests/cpu_proofs/verify_math.c): c/255 for all 256 codes, c/65535 for all 65536 codes and h/60 for
every float in [2^-20, 720] must equal IEEE division bit for bit; the float -> code conversion of
colorlut (RZ add of 0.5, RD add of 2^23, low mantissa bits) must equal round-half-away for every
float in [0, 1] at 8 and 16 bits; the sector / triangle-wave form of HSV -> RGB must equal the
reference's fmod form and `<=` ladder for every float h/60 in [0, 6]."""
import os
import subprocess


def test_division_shortcuts_proven(tmp_path):
    src = os.path.join(os.path.dirname(__file__), "cpu_proofs", "verify_math.c")
    exe = tmp_path / "verify_math"
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-o", str(exe), src, "-lm"], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout
    assert "ALL PROVEN" in out.stdout
    # the constants proven are the ones the kernels use
    hdr = open(os.path.join(os.path.dirname(__file__), "..", "gst-plugins-rs_b200", "csrc",
                            "vf_math.cuh")).read()
    for tok in ("0x1.010102p-8f", "-0x1.fdfdfep-33f", "0x1.0001p-16f", "0x1.0001p-48f",
                "0x1.111112p-6f", "-0x1.dddddep-31f"):
        assert tok in hdr
    for line in ("K255  hi=0x1.010102p-8 lo=-0x1.fdfdfep-33", "K65535 hi=0x1.0001p-16 lo=0x1.0001p-48",
                 "K60  hi=0x1.111112p-6 lo=-0x1.dddddep-31", "q60b (two-term): 0 failures",
                 "round8: 1065353217 values, 0 failures", "round16: 1065353217 values, 0 failures",
                 "wave: 1086324737 values, 0 failures; arm: 0 failures"):
        assert line in out.stdout


def test_microbenchmarks_build_and_self_check(tmp_path):
    """tools/microbench/*.cu are evidence behind DESIGN.md (§4, §7, §13): they must keep compiling for
    sm_100a, and smem_lut's host-side reference (exact /255 split, identity LUT) must hold — it
    runs before the first CUDA call, so it can be checked without a device."""
    import glob
    import shutil
    if not shutil.which("nvcc"):
        import pytest
        pytest.skip("nvcc not available")
    mb = os.path.join(os.path.dirname(__file__), "..", "tools", "microbench")
    from concurrent.futures import ThreadPoolExecutor

    def compile_one(src):
        exe = tmp_path / os.path.basename(src)[:-3]
        subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17",
                        "-fmad=false", "-o", str(exe), src], check=True, timeout=900)

    with ThreadPoolExecutor(max_workers=4) as pool:   # the two that include csrc/ take a minute each
        list(pool.map(compile_one, sorted(glob.glob(os.path.join(mb, "*.cu")))))
    out = subprocess.run([str(tmp_path / "smem_lut")], capture_output=True, text=True, timeout=120)
    assert "host self-check ok" in out.stdout, out.stdout + out.stderr
