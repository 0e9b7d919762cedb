"""e2e host path: slots x chunk sweep with SM / memory clocks sampled during the run (dev aid)."""
import subprocess
import os
import sys
os.environ["B200VF_ALLOW_DEBUG_MODES"] = "1"  # modes 1-3 skip the kernels (wrong pixels): analysis only
import threading
import time
import torch
sys.path.insert(0, ".")
import gst_plugins_rs_b200 as g
from gst_plugins_rs_b200 import frames
from gst_plugins_rs_b200.api import frame_array, frame_of

w, h, nb = 3840, 2160, 8
fb = w * h * 4
ctx = g.Context(0)
ctx.set_lut_from_cube(g.parse_cube(frames.cube_text_3d(65)))
src = frames.frame_grad(w, h).reshape(-1)
hin = [torch.from_numpy(src.copy()).pin_memory() for _ in range(nb)]
hout = [torch.empty_like(t).pin_memory() for t in hin]
fi = frame_array([frame_of(t, w, h, "RGBA") for t in hin])
fo = frame_array([frame_of(t, w, h, "RGBA") for t in hout])
clk = []


def sample():
    out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,pstate,power.draw", "--format=csv,noheader"],
                         capture_output=True, text=True).stdout.strip()
    clk.append(out)


for mode in (0,):
    ctx.set_option("host.dbg_mode", mode)
    for slots in (3, 4, 6):
        ctx.set_option("host.slots", slots)
        for chunk in (0, 4 << 20, 8 << 20, 16 << 20, 34 << 20):
            ctx.set_option("host.chunk_bytes", chunk)
            for _ in range(2):
                ctx.colorlut_batch(fi, fo)
            c0 = ctx.get_option("host.dbg_chunks")
            th = threading.Timer(0.03, sample)
            th.start()
            t0 = time.perf_counter()
            iters = 12
            for _ in range(iters):
                ctx.colorlut_batch(fi, fo)
            dt = time.perf_counter() - t0
            th.join()
            print("mode %d slots %d chunk %2d MB: %5.0f frames/s %5.1f GB/s each way, %3d chunks/call  clocks[%s]" %
                  (mode, slots, chunk >> 20, nb * iters / dt, nb * iters * fb / dt / 1e9,
                   (ctx.get_option("host.dbg_chunks") - c0) // iters, clk[-1] if clk else ""))
