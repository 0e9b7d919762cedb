"""BASELINE.json configs[0] side by side on this box: the C++ element pipeline
(examples/cfg1_pipeline, system-memory frames through libb200vf.so) vs the CPU restatement of the
reference on ONE thread (the reference element processes one frame at a time on its streaming
thread).  300 SMPTE-like 1920x1080 RGBA buffers, 33^3 LUT, trilinear."""
import json
import os
import subprocess
import sys
import tempfile
import time

sys.path.insert(0, ".")
import numpy as np

import oracle
from gst_plugins_rs_b200 import frames


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    text = frames.cube_text_3d(33)
    with tempfile.NamedTemporaryFile("w", suffix=".cube", delete=False) as f:
        f.write(text)
    out = subprocess.run(["examples/cfg1_pipeline", f.name, str(n)], capture_output=True, text=True)
    gpu = json.loads(out.stdout.strip().splitlines()[-1])
    lut = oracle.Lut(text=text)
    src = frames.frame_bars(1920, 1080).reshape(-1)
    dst = np.empty_like(src)
    k = min(n, 30)
    t0 = time.perf_counter()
    for _ in range(k):
        oracle.colorlut(lut, src, 1920, 1080, dst=dst)
    cpu_fps = k / (time.perf_counter() - t0)
    print(json.dumps({"cfg1_gpu_pipeline": gpu, "cfg1_cpu_reference_port_1thread_fps": cpu_fps,
                      "cpu_frames_timed": k, "speedup": gpu["frames_per_s"] / cpu_fps}))
    os.unlink(f.name)


if __name__ == "__main__":
    main()
