"""bench.py contract (CPU-checkable part): the reference arm prints ONE JSON line with the keys the
driver reads, for the same metric/config vocabulary as the B200 arm."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True,
                         text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stdout.strip().splitlines()


def test_reference_arm_json_line():
    lines = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--workload",
                 "hsvfilter_1080p")
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["vs_baseline"] is None
    assert d["config"]["workload"] == "hsvfilter_1080p" and d["data"] == "synthetic"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_do_no_work():
    """Under torchrun only rank 0 runs the CPU arm; the other ranks exit 0 silently."""
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    assert _run("--impl", "reference", "--steps", "1", "--workload", "hsvfilter_1080p", env=env) == []


def test_workload_table_names_baseline_configs():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.HEADLINE == "colorlut65_4k"

    def key(name):
        s = bench.WORKLOADS[name]
        return (s["element"], s["width"], s["height"], s["lut"], s["in_fmt"], s["out_fmt"])
    assert key("colorlut65_4k") == ("colorlut", 3840, 2160, 65, "RGBA", "RGBA")        # configs[2]
    assert key("hsvfilter_1080p") == ("hsvfilter", 1920, 1080, 0, "RGBA", "RGBA")      # configs[1]
    assert key("hsvdetector_4k") == ("hsvdetector", 3840, 2160, 0, "BGRx", "RGBA")     # configs[3]
    assert key("chain33_8k") == ("chain", 7680, 4320, 33, "RGBA", "RGBA")              # configs[4]
    assert key("colorlut33_1080p") == ("colorlut", 1920, 1080, 33, "RGBA", "RGBA")     # configs[0]
    assert bench.WORKLOADS["colorlut65_4k"]["options"] == {}   # the headline runs the library's defaults


def test_workload_variants_map_to_reference_elements():
    """Table / interpolation / format variants are bench-side labels; the element underneath is
    one of the reference's three (or the chain), and the CPU arm runs exactly that."""
    sys.path.insert(0, ROOT)
    import bench
    for name, s in bench.WORKLOADS.items():
        assert s["element"] in ("colorlut", "colorlut_convert", "hsvfilter", "hsvdetector", "chain", "pipelines"), name
        assert s["width"] % 4 == 0 and s["lut"] in (0, 33, 65, 1024)
        assert set(s["options"]) <= {"lut.path", "lut.interpolation", "hsv.path", "tables.share"}
        # algorithmic bytes of SURVEY.md §8(d): bytes read + bytes written per pixel
        want = {"RGBA": 8, "BGRx->RGBA": 8, "BGRx->BGRA": 8, "RGBA64_LE": 16, "RGB": 6, "RGB->RGBA": 7}[
            bench.workload_config(name, "grad")["format"]]
        assert s["bytes_per_pixel"] == want * (2 if s["element"] == "pipelines" else 1), name
    assert not bench.HEADLINE.endswith(("_interp", "_direct", "_compute", "_tetrahedral"))


def test_both_arms_describe_the_workload_identically():
    """`config` names the workload only; what differs between the arms (frames per step, kernels,
    parallelism) lives under `run`, so the driver's same-config check compares like with like."""
    sys.path.insert(0, ROOT)
    import bench
    d = json.loads(_run("--impl", "reference", "--steps", "1", "--warmup", "0", "--workload",
                        "hsvfilter_1080p", "--content", "noise")[0])
    assert d["config"] == bench.workload_config("hsvfilter_1080p", "noise")
    assert set(d["config"]) == {"workload", "element", "width", "height", "format", "lut", "content",
                                "pipelines"}
    assert "frames_per_step" in d["run"]


def test_workload_frame_wide_is_the_same_colour_at_16_bit():
    sys.path.insert(0, ROOT)
    import numpy as np
    import bench
    f8 = bench.workload_frame("grad", 64, 8, 0)
    f16 = bench.workload_frame("grad", 64, 8, 0, wide=True)
    assert f8.dtype == np.uint8 and f16.dtype == np.uint8 and f16.size == 2 * f8.size
    assert np.array_equal(f16.view("<u2"), f8.astype(np.uint16) * 257)


def test_reference_arm_runs_a_colorlut_workload():
    d = json.loads(_run("--impl", "reference", "--steps", "1", "--warmup", "0", "--workload",
                        "colorlut33_1080p")[0])
    assert d["config"]["element"] == "colorlut" and d["value"] > 0
