#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 colour-transform path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload NAME] [--content bars|grad|noise|rand] [--batch B] [--no-extras]
                    [--single-process]

One "step" = one pass of the hot path over one batch of B synthetic frames (B = 512 4K frames for
the headline, i.e. 8 launches of 64 frames and ~5 ms of device time per step, so that the K = 20
timed steps cover >= 100 ms).
Headline workload (N=1): `colorlut` 65^3 LUT, trilinear (the reference's only 3D mode,
SURVEY.md F1), 3840x2160 RGBA — BASELINE.json configs[2] with the parity-checked
interpolation.  `value` = 4K RGBA frames/s with frames resident in HBM; `e2e` = the same
metric through the C ABI with pinned HOST frames (H2D + kernel + D2H inside the timed
region); `roofline` = algorithmic bytes (8*W*H per frame) / CUDA-event time against the
measured HBM copy peak in MEASURED_PEAKS.json.  `content_classes` repeats the headline workload
on all four content classes of SURVEY.md §8(d) (+ the camera-like `noise` class) and
`worst_class` names the slowest — table gathers are content-sensitive, so the headline never
stands without it.  The other elements/configs are measured the same way and reported under
"workloads" (each a parity-test case, not the headline).

`--impl reference` times the CPU restatement of the reference (oracle/, the Rust
toolchain being absent) on the box's host cores, frame-parallel on all of them.

Multi-GPU: frames are independent, so each GPU processes its own batches with no data-path
collective ("weak" scaling).  Under torchrun (one rank per GPU) the timed region is bracketed by
barrier + synchronize and the max over ranks is taken; `--single-process` drives the same N GPUs
from ONE process through the library's group dispatcher (b200vf_group_*: one host thread and three
streams per device, frame i -> device i mod N) and takes the max over the devices' CUDA events.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG2 = (37.5, 1.2, 0.05, 0.9, 0.02)            # hsvfilter settings of SURVEY.md §8(d) cfg2
DET_CFG4 = (120.0, 30.0, 0.6, 0.4, 0.6, 0.4)   # hsvdetector settings of cfg4
CONTENT_CLASSES = ("bars", "grad", "noise", "rand")


def W(element, width, height, lut=0, lut_kind="3d", in_fmt="RGBA", out_fmt=None, options=None,
      contents=None, pipelines=1):
    """One workload: a reference element (or the chain) + geometry + formats + library options."""
    bpp = {"RGB": 3, "BGR": 3, "RGBA64_LE": 8, "RGBA64_BE": 8}
    out_fmt = out_fmt or in_fmt
    return {"element": element, "width": width, "height": height, "lut": lut, "lut_kind": lut_kind,
            "in_fmt": in_fmt, "out_fmt": out_fmt, "options": options or {},
            "contents": contents, "pipelines": pipelines,
            # algorithmic bytes per pixel: every pixel read once and written once;
            # two element passes for the unfused pipelines workload
            "bytes_per_pixel": (bpp.get(in_fmt, 4) + bpp.get(out_fmt, 4)) * (2 if element == "pipelines" else 1)}


ALL = CONTENT_CLASSES
WORKLOADS = {
    "colorlut65_4k": W("colorlut", 3840, 2160, 65, contents=ALL),                     # configs[2]
    "colorlut33_4k": W("colorlut", 3840, 2160, 33),
    "hsvfilter_4k": W("hsvfilter", 3840, 2160, contents=ALL),
    "hsvfilter_1080p": W("hsvfilter", 1920, 1080),                                    # configs[1]
    # "hsv.path"=1: always the compute kernels (the reference's f32 sequence per pixel); the default
    # (auto) serves from the function table whenever that measures faster on the stream's frames
    "hsvfilter_4k_compute": W("hsvfilter", 3840, 2160, options={"hsv.path": 1}),
    "hsvdetector_4k_compute": W("hsvdetector", 3840, 2160, in_fmt="BGRx", out_fmt="RGBA",
                                options={"hsv.path": 1}),
    "hsvdetector_4k": W("hsvdetector", 3840, 2160, in_fmt="BGRx", out_fmt="RGBA", contents=ALL),  # configs[3]
    "chain33_8k": W("chain", 7680, 4320, 33),                                         # configs[4]
    "colorlut33_1080p": W("colorlut", 1920, 1080, 33),                                # configs[0]
    # "lut.path"=3: the interpolating kernel (R- and G-resampled table, z-lerp per pixel) that serves
    # when the 64 MiB baked table cannot be allocated — reported next to the default so both designs
    # stay measured; "lut.path"=1: the direct 8-corner kernel the auto policy falls back to
    "colorlut65_4k_interp": W("colorlut", 3840, 2160, 65, options={"lut.path": 3}, contents=ALL),
    "colorlut65_4k_direct": W("colorlut", 3840, 2160, 65, options={"lut.path": 1}),
    # EXTENSION modes ("lut.interpolation" = 1 / 2): BASELINE.json's configs name tetrahedral, the
    # reference implements trilinear only — parity is against the oracle's own
    # definition, so these are reported next to the headline, never as it
    "colorlut65_4k_tetrahedral": W("colorlut", 3840, 2160, 65, options={"lut.interpolation": 1}, contents=ALL),
    "colorlut65_4k_tetrahedral_direct": W("colorlut", 3840, 2160, 65,
                                          options={"lut.interpolation": 1, "lut.path": 1}),
    "colorlut65_4k_nearest": W("colorlut", 3840, 2160, 65, options={"lut.interpolation": 2}),
    # RGBA64_LE, the first format in the reference element's caps: 16 B/pixel, direct 8-corner path
    "colorlut33_4k_rgba64": W("colorlut", 3840, 2160, 33, in_fmt="RGBA64_LE", contents=ALL),
    # the remaining formats of the elements' caps
    "hsvfilter_4k_rgb": W("hsvfilter", 3840, 2160, in_fmt="RGB"),                     # 3-byte pixels, 6 B/px
    "hsvfilter_4k_rgb_compute": W("hsvfilter", 3840, 2160, in_fmt="RGB", options={"hsv.path": 1}),
    "hsvdetector_4k_rgb": W("hsvdetector", 3840, 2160, in_fmt="RGB", out_fmt="RGBA"),  # 7 B/px
    "colorlut1d_4k": W("colorlut", 3840, 2160, 1024, lut_kind="1d"),
    "colorlut1d_4k_rgba64": W("colorlut", 3840, 2160, 1024, lut_kind="1d", in_fmt="RGBA64_LE"),
    # `videoconvert ! colorlut ! videoconvert` (the reference's own example pipeline, colorlut/imp.rs:17-19)
    # as ONE pass: BGRx in, BGRA out, the two conversions folded into the gather (extension; the CPU
    # arm times the colorlut step alone, i.e. less work than the reference pipeline does)
    "colorlut33_4k_convert_bgrx": W("colorlut_convert", 3840, 2160, 33, in_fmt="BGRx", out_fmt="BGRA"),
    # four pipelines `colorlut(33^3) ! hsvfilter` on ONE GPU, each element its own context (8
    # contexts): what a multi-stream application looks like to the device-wide table cache
    # (DESIGN.md §14).  "tables.share"=0 gives every context private tables, as in round 1.
    "pipelines4_lut33_hsv_4k": W("pipelines", 3840, 2160, 33, pipelines=4, contents=("grad", "noise", "rand")),
    "pipelines4_lut33_hsv_4k_private": W("pipelines", 3840, 2160, 33, pipelines=4,
                                         options={"tables.share": 0}, contents=("grad", "noise", "rand")),
}
HEADLINE = "colorlut65_4k"
PROFILE_MODE = False
HSV_PATH = 0  # "hsv.path" of the default workloads; --hsv-path 2 pins the table kernel for ncu captures
EXTRA_OPTIONS = {}  # --option key=value: library options applied on top of the workload's (ncu captures)


def workload_config(name, content):
    """The `config` object of the JSON line: what is computed, on what — identical for both arms."""
    s = WORKLOADS[name]
    lut = None
    if s["lut"]:
        interp = {0: "trilinear", 1: "tetrahedral (extension)", 2: "nearest (extension)"}[
            s["options"].get("lut.interpolation", 0)]
        lut = (f"1D size {s['lut']} synthetic .cube" if s["lut_kind"] == "1d"
               else f"{s['lut']}^3 synthetic .cube, {interp}")
    return {"workload": name, "element": s["element"], "width": s["width"], "height": s["height"],
            "format": s["in_fmt"] if s["in_fmt"] == s["out_fmt"] else f"{s['in_fmt']}->{s['out_fmt']}",
            "lut": lut, "content": content, "pipelines": s["pipelines"]}


def metric_name(name):
    return ("4K RGBA frames/sec (colorlut 65^3 trilinear; hsvfilter under workloads)"
            if name == HEADLINE else f"frames/sec ({name})")


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms",
                 "20", "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([time.perf_counter()] + [c.strip() for c in line.split(",")][1:])

    def stop(self, window=None):
        """window = (t0, t1) in time.perf_counter() terms: keep only the samples taken while the
        GPU was under load (nvidia-smi is started early because it needs ~0.3 s to come up)."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            if window and not (window[0] <= r[0] <= window[1] + 0.02):
                continue
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 4), ("hw_thermal_slowdown", 5),
                              ("sw_thermal_slowdown", 6), ("sw_power_cap", 7)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "samples": len(sm),
                "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------
def workload_frame(content, w, h, index, wide=False, bpp3=False):
    """One synthetic frame of a content class as flat bytes; `wide` = RGBA64_LE (each 8-bit code c
    becomes the 16-bit code 257*c, i.e. the same colour at full 16-bit scale); `bpp3` = the same
    colours as packed 3-byte pixels."""
    from gst_plugins_rs_b200 import frames
    f = frames.frame_of_class(content, w, h, index).reshape(-1)
    if wide:
        f = (f.astype("<u2") * 257).view(np.uint8).reshape(-1)
    if bpp3:
        f = np.ascontiguousarray(f.reshape(-1, 4)[:, :3]).reshape(-1)
    return f


def lut_text(spec):
    from gst_plugins_rs_b200 import frames
    return frames.cube_text_1d(spec["lut"]) if spec["lut_kind"] == "1d" else frames.cube_text_3d(spec["lut"])


class Engine:
    """What runs the steps: one context (one GPU per process), or a group of member contexts over
    several GPUs driven from this process (`--single-process`)."""

    def __init__(self, g, devices, single_process):
        import torch
        self.g, self.devices, self.group = g, devices, None
        if single_process:
            self.group = g.Group(devices)
            self.ctxs = [self.group.member(m) for m in range(len(devices))]
            self.streams = [torch.cuda.ExternalStream(c.get_stream(), device=f"cuda:{d}")
                            for c, d in zip(self.ctxs, devices)]
            self.api = self.group
        else:
            ctx = g.Context(devices[0])
            ctx.set_stream(torch.cuda.current_stream().cuda_stream)
            self.ctxs = [ctx]
            self.streams = [torch.cuda.current_stream()]
            self.api = ctx

    def set_option(self, key, value):
        self.api.set_option(key, value)

    def synchronize(self):
        import torch
        for d in self.devices:
            torch.cuda.synchronize(d)

    def reset_stats(self):
        for c in self.ctxs:
            c.reset_stats()

    def stats(self):
        tot = {"kernel_launches": 0, "frames": 0, "h2d_bytes": 0, "d2h_bytes": 0}
        for c in self.ctxs:
            for k, v in c.stats().items():
                tot[k] += v
        return tot


class Runner:
    """Holds one workload's device/host buffers and the closures that run one step."""

    def __init__(self, g, eng, name, content, batch, rank):
        import torch
        from gst_plugins_rs_b200.api import frame_array, frame_of
        self.name, self.eng, self.g = name, eng, g
        s = self.spec = WORKLOADS[name]
        self.elem, self.w, self.h = s["element"], s["width"], s["height"]
        w, h = self.w, self.h
        opts = {"lut.path": 0, "lut.interpolation": 0, "hsv.path": HSV_PATH, "tables.share": 1}
        opts.update(s["options"])
        opts.update(EXTRA_OPTIONS)
        self.n_dev = len(eng.devices)
        self.batch = batch * self.n_dev  # frame i lives on (and is processed by) device i mod N
        self.in_fmt, self.out_fmt = s["in_fmt"], s["out_fmt"]
        self.bytes_per_frame = s["bytes_per_pixel"] * w * h
        self.pipes = []
        if self.elem == "pipelines":  # each element of each pipeline is its own context, as in GStreamer
            for _ in range(s["pipelines"]):
                lut_ctx, hsv_ctx = g.Context(eng.devices[0]), g.Context(eng.devices[0])
                for c in (lut_ctx, hsv_ctx):
                    c.set_stream(torch.cuda.current_stream().cuda_stream)
                    for k, v in opts.items():
                        c.set_option(k, v)
                lut_ctx.set_lut_from_cube(g.parse_cube(lut_text(s)))
                self.pipes.append((lut_ctx, hsv_ctx))
        else:
            for k, v in opts.items():
                eng.set_option(k, v)
            if s["lut"]:
                eng.api.set_lut_from_cube(g.parse_cube(lut_text(s)))
        # distinct synthetic frames; the batch working set (in + out) exceeds the 126 MB L2
        uniq = min(self.batch, 4)
        wide, bpp3 = self.in_fmt.startswith("RGBA64"), self.in_fmt in ("RGB", "BGR")
        host = [workload_frame(content, w, h, rank * 1000 + i, wide, bpp3) for i in range(uniq)]
        self.src_np = host
        out_bytes = w * h * {"RGB": 3, "BGR": 3, "RGBA64_LE": 8, "RGBA64_BE": 8}.get(self.out_fmt, 4)
        self.d_in, self.d_out = [], []
        seeds = {}  # one upload per (device, distinct frame); every further copy is a device clone
        for i in range(self.batch):
            dev = eng.devices[i % self.n_dev]
            key = (dev, i % uniq)
            if key not in seeds:
                seeds[key] = torch.from_numpy(host[i % uniq]).to(f"cuda:{dev}")
                self.d_in.append(seeds[key])
            else:
                self.d_in.append(seeds[key].clone())  # no aliasing: the working set must exceed L2
            self.d_out.append(torch.empty(out_bytes, dtype=torch.uint8, device=f"cuda:{dev}"))
        self.fin = frame_array([frame_of(t, w, h, self.in_fmt) for t in self.d_in])
        self.fout = frame_array([frame_of(t, w, h, self.out_fmt) for t in self.d_out])
        self.fscratch = frame_array([frame_of(t, w, h, self.in_fmt) for t in self.d_out])
        self.hp = g.HsvFilterParams(*CFG2)
        self.dp = g.HsvDetectorParams(*DET_CFG4)
        self.h_in = self.h_out = None
        # hsvfilter works in place: a buffer that is filtered again and again stops being a frame of
        # its content class (CFG2 drives every pixel towards s = 1, v = 0.2).  The kernel that serves
        # from the function table is content-sensitive, so every TIMED step gets buffers holding
        # pristine content (`fresh` rings, prepared before the timed region), and warm-up / soak
        # steps restore theirs from `d_in` first (device copy, outside any timed interval).
        self.inplace = self.elem == "hsvfilter"
        self.fresh, self.cursor = [], 0
        if self.pipes:  # every pipeline filters its own share of the batch
            per = self.batch // len(self.pipes)
            assert per * len(self.pipes) == self.batch
            self.pipe_frames = [(frame_array(list(self.fin[i * per:(i + 1) * per])),
                                 frame_array(list(self.fout[i * per:(i + 1) * per])))
                                for i in range(len(self.pipes))]

    def close(self):
        for a, b in self.pipes:
            a.close(), b.close()
        self.pipes = []

    def prepare_timed(self, steps):
        """Buffers for `steps` timed in-place steps (no-op for out-of-place elements)."""
        from gst_plugins_rs_b200.api import frame_array, frame_of
        if not self.inplace:
            return steps
        self.fresh, self.cursor = [], 0
        budget = 40 << 30
        ring_bytes = sum(t.numel() for t in self.d_in)
        rings = max(1, min(steps, budget // ring_bytes))
        for _ in range(rings):
            bufs = [t.clone() for t in self.d_in]
            self.fresh.append((bufs, frame_array([frame_of(t, self.w, self.h, self.in_fmt) for t in bufs])))
        return rings  # every timed step filters pristine content: no ring is used twice

    def restore(self):
        """Pristine content back into the scratch buffers of an in-place element."""
        if self.inplace:
            for dst, src in zip(self.d_out, self.d_in):
                dst.copy_(src)

    def step_device(self, timed=False):
        c = self.eng.api
        if self.elem == "colorlut":
            c.colorlut_batch(self.fin, self.fout)
        elif self.elem == "colorlut_convert":
            c.colorlut_convert_batch(self.fin, self.fout)
        elif self.elem == "hsvfilter":  # in place, like the element
            if timed and self.fresh:
                c.hsvfilter_batch(self.fresh[self.cursor % len(self.fresh)][1], self.hp)
                self.cursor += 1
            else:
                c.hsvfilter_batch(self.fscratch, self.hp)
        elif self.elem == "hsvdetector":
            c.hsvdetector_batch(self.fin, self.fout, self.dp)
        elif self.elem == "pipelines":
            for (lut_ctx, hsv_ctx), (a, b) in zip(self.pipes, self.pipe_frames):
                lut_ctx.colorlut_batch(a, b)
                hsv_ctx.hsvfilter_batch(b, self.hp)
        else:
            c.chain_lut_hsv_batch(self.fin, self.fout, self.hp)

    def stats(self):
        if not self.pipes:
            return self.eng.stats()
        tot = {"kernel_launches": 0, "frames": 0, "h2d_bytes": 0, "d2h_bytes": 0}
        for pair in self.pipes:
            for c in pair:
                for k, v in c.stats().items():
                    tot[k] += v
        return tot

    def reset_stats(self):
        self.eng.reset_stats()
        for pair in self.pipes:
            for c in pair:
                c.reset_stats()

    def prepare_host(self, e2e_batch):
        import torch
        from gst_plugins_rs_b200.api import frame_array, frame_of
        w, h = self.w, self.h
        self.e2e_batch = e2e_batch * self.n_dev
        self.h_in = [torch.from_numpy(self.src_np[i % len(self.src_np)].copy()).pin_memory()
                     for i in range(self.e2e_batch)]
        self.h_out = [torch.empty(self.d_out[0].numel(), dtype=torch.uint8).pin_memory() for _ in self.h_in]
        self.hfin = frame_array([frame_of(t, w, h, self.in_fmt) for t in self.h_in])
        self.hfout = frame_array([frame_of(t, w, h, self.out_fmt) for t in self.h_out])
        self.hsets = [(self.hfin, self.hfout)]

    def prepare_host_queued(self):
        """A second set of pinned frames: with calls in flight ("host.async") the frames of the
        call still running belong to the library, the caller fills / reads the other set."""
        import torch
        from gst_plugins_rs_b200.api import frame_array, frame_of
        w, h = self.w, self.h
        self.h_in2 = [t.clone().pin_memory() for t in self.h_in]
        self.h_out2 = [torch.empty_like(t).pin_memory() for t in self.h_out]
        self.hsets = [(self.hfin, self.hfout),
                      (frame_array([frame_of(t, w, h, self.in_fmt) for t in self.h_in2]),
                       frame_array([frame_of(t, w, h, self.out_fmt) for t in self.h_out2]))]

    def step_host(self, which=0):
        """The call a pipeline makes with system-memory buffers: complete on return (or, with
        "host.async", when the caller waits for its ticket)."""
        c = self.eng.api
        fin, fout = self.hsets[which]
        if self.elem == "colorlut":
            c.colorlut_batch(fin, fout)
        elif self.elem == "colorlut_convert":
            c.colorlut_convert_batch(fin, fout)
        elif self.elem == "hsvfilter":
            c.hsvfilter_batch(fin, self.hp)
        elif self.elem == "hsvdetector":
            c.hsvdetector_batch(fin, fout, self.dp)
        else:
            c.chain_lut_hsv_batch(fin, fout, self.hp)


LUT_KERNELS = {0: "direct 8-corner interpolation", 1: "R-resampled table + 2 lerps", 2: "1D",
               3: "RG-resampled table + z-lerp", 5: "tetrahedral", 6: "nearest",
               7: "16-bit delta-table op (x-differences precomputed, packed f32x2 lerps, constant strides, "
                  "FRND/F2I coordinates)",
               4: "table baked to 8-bit resolution by the direct kernel, 4x4x2 colour blocks per line, "
                  "2-D tile traversal (one gather per pixel)"}


def kernel_of(r):
    """Which kernel served the last steps — the library picks."""
    if r.pipes:
        lut_ctx, hsv_ctx = r.pipes[0]
        return ("colorlut: " + LUT_KERNELS.get(lut_ctx.get_option("lut.path_active"), "?") + " | hsvfilter: " +
                ("function table" if hsv_ctx.get_option("hsv.table_active") else "compute kernels"))
    ctx = r.eng.ctxs[0]
    if r.elem in ("colorlut", "colorlut_convert"):
        return r.elem + ": " + LUT_KERNELS.get(ctx.get_option("lut.path_active"), "?")
    return r.elem + (": function table filled by the compute kernels (one gather per pixel, blocked layout, "
                     "2-D tiles)" if ctx.get_option("hsv.table_active") else
                     ": compute kernels (the reference's f32 sequence per pixel)")


def bind_to_gpu_numa_node(local):
    """Pin this rank's threads to the CPUs next to its GPU (sysfs local_cpulist) before any host
    buffer is allocated, so pinned frames land on the GPU's own NUMA node (first touch) and the
    PCIe DMA of one rank does not cross the socket interconnect.  Best effort."""
    try:
        import torch
        p = torch.cuda.get_device_properties(local)
        dev = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        cpulist = open(f"/sys/bus/pci/devices/{dev}/local_cpulist").read().strip()
        cpus = set()
        for part in cpulist.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return cpulist
    except Exception:
        pass
    return None


def dist_setup(single_process):
    import torch
    from gst_plugins_rs_b200 import sharding
    rank, local, world = sharding.world()
    if single_process:
        return 0, 0, 1, False
    torch.cuda.set_device(local)
    if world > 1:
        bind_to_gpu_numa_node(local)
    use_dist = sharding.init_process_group("nccl", torch.device("cuda", local))
    return rank, local, world, use_dist


def barrier_sync(eng, use_dist):
    from gst_plugins_rs_b200 import sharding
    if use_dist:
        sharding.barrier()
    eng.synchronize()


def max_over_ranks(ms, use_dist):
    from gst_plugins_rs_b200 import sharding
    return sharding.max_over_ranks(ms, "cuda") if use_dist else ms


def time_device(r, steps, warmup, use_dist, soak_s=0.3):
    """K steps timed with CUDA events on the launching stream(s), barrier+sync on both sides.
    After the W warm-up steps the kernel keeps running for `soak_s` seconds, so the timed steps see
    the steady-state clocks (on these 1 kW parts: the power-capped ones) that nvidia-smi samples.
    Returns (ms, launches, steps actually timed)."""
    import torch
    eng = r.eng
    steps = r.prepare_timed(steps)
    for _ in range(warmup):
        r.restore()
        r.step_device()
    eng.synchronize()
    t0 = time.perf_counter()
    r.load_window = [t0 + 0.1, t0]  # clocks are sampled from 0.1 s into the soak to the end of the timed steps
    while not PROFILE_MODE and time.perf_counter() - t0 < soak_s:
        for _ in range(4):
            r.restore()
            r.step_device()
        eng.synchronize()
    barrier_sync(eng, use_dist)
    r.reset_stats()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in eng.streams]
    for (e0, _), s in zip(ev, eng.streams):
        e0.record(s)
    for _ in range(steps):
        r.step_device(timed=True)
    for (_, e1), s in zip(ev, eng.streams):
        e1.record(s)
    barrier_sync(eng, use_dist)
    r.load_window[1] = time.perf_counter()
    ms = max(e0.elapsed_time(e1) for e0, e1 in ev)  # max over this process's devices …
    launches = r.stats()["kernel_launches"]
    return max_over_ranks(ms, use_dist), launches, steps  # … and over the ranks


def pcie_bidir_peak_gbs():
    """Pinned cudaMemcpyAsync H2D and D2H running concurrently on two streams (GB/s each way):
    the ceiling of the e2e path on this box."""
    import torch
    n = 128 << 20
    h1 = torch.empty(n, dtype=torch.uint8).pin_memory()
    h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
    d1 = torch.empty(n, dtype=torch.uint8, device="cuda")
    d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def both():
        with torch.cuda.stream(s1):
            d1.copy_(h1, non_blocking=True)
        with torch.cuda.stream(s2):
            h2.copy_(d2, non_blocking=True)
    both()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(6):
        both()
    torch.cuda.synchronize()
    return n * 6 / (time.perf_counter() - t0) / 1e9


def time_host(r, steps, warmup, use_dist):
    """End to end through the C ABI with pinned host frames; wall clock around synchronous
    calls (each call returns only when its D2H has landed), bracketed like the device run."""
    for _ in range(max(1, min(warmup, 2))):
        r.step_host()
    barrier_sync(r.eng, use_dist)
    r.reset_stats()
    t0 = time.perf_counter()
    for _ in range(steps):
        r.step_host()
    barrier_sync(r.eng, use_dist)
    ms = (time.perf_counter() - t0) * 1e3
    st = r.stats()
    return max_over_ranks(ms, use_dist), st


def time_host_queued(r, steps, warmup, use_dist):
    """The same calls with one call of latency — what an element that holds one buffer back does
    (BaseTransform's submit_input_buffer / generate_output): "host.async" = 1, call k is queued,
    then the caller waits for call k-1's ticket; two sets of pinned frames alternate.  Every
    step's H2D and D2H are inside the timed region; the last call is waited for before it ends."""
    eng = r.eng
    r.prepare_host_queued()
    eng.set_option("host.async", 1)

    def run(n):
        prev = None
        for k in range(n):
            r.step_host(k & 1)
            cur = [c.host_ticket() for c in eng.ctxs]
            if prev is not None:
                for c, t in zip(eng.ctxs, prev):
                    c.host_wait(t)
            prev = cur
        for c, t in zip(eng.ctxs, prev):
            c.host_wait(t)

    run(max(2, min(warmup, 2)))
    barrier_sync(eng, use_dist)
    t0 = time.perf_counter()
    run(steps)
    barrier_sync(eng, use_dist)
    ms = (time.perf_counter() - t0) * 1e3
    eng.set_option("host.async", 0)
    return max_over_ranks(ms, use_dist)


def default_batch(name):
    """Frames per step and per GPU: ~5 ms of device time per step for the headline (512 4K frames
    = 8 launches of 64); one full launch for the other workloads (64 frames at 4K and 1080p, 16 at
    8K: 0.5-4 GB working sets, 10-40 ms of timed device work over the 20 steps) — always larger
    than the 126 MB L2."""
    s = WORKLOADS[name]
    px = s["width"] * s["height"]
    if name == HEADLINE:
        return 512
    return max(4, min(64, (4 << 30) // (8 * px)))


def measure_workload(g, eng, name, content, batch, steps, warmup, use_dist, rank, world, peak):
    """Device-resident measurement of one workload / content class → result dict (+ the Runner)."""
    r = Runner(g, eng, name, content, batch, rank)
    ms, launches, steps = time_device(r, steps, warmup, use_dist)
    n_dev = len(eng.devices)
    frames_total = r.batch * steps * world
    gbs_per_gpu = r.bytes_per_frame * r.batch * steps / (ms / 1e3) / 1e9 / n_dev
    res = {"frames_per_s": frames_total / (ms / 1e3), "gbs_per_gpu": gbs_per_gpu,
           "frac_of_hbm_peak": gbs_per_gpu / peak, "frames_per_step": r.batch, "steps": steps,
           "timed_ms": ms, "launches": int(launches), "kernel": kernel_of(r)}
    return res, r, ms, launches, steps


def run_b200(args):
    # stdout carries exactly one JSON line: libraries that print there (NCCL's version banner under
    # torchrun) are sent to stderr for the duration of the run
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    import torch
    import gst_plugins_rs_b200 as g
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 arm has no CPU fallback")
    rank, local, world, use_dist = dist_setup(args.single_process)
    peak, peak_src = load_peaks()
    devices = list(range(args.gpus)) if args.single_process else [local]
    n_gpus = args.gpus if args.single_process else world
    eng = Engine(g, devices, args.single_process)

    sampler = ClockSampler(devices[0]) if rank == 0 else None
    if sampler:
        sampler.start()
    name = args.workload
    batch = args.batch or default_batch(name)
    res, r, ms, launches, steps = measure_workload(g, eng, name, args.content, batch, args.steps,
                                                   args.warmup, use_dist, rank, world, peak)
    clocks = sampler.stop(tuple(r.load_window)) if sampler else None
    kernel_note = res["kernel"]
    value = res["frames_per_s"]
    kernel_ms = ms / max(1, launches) * len(devices)  # launches are summed over this process's devices
    frames_per_launch = r.batch * steps / max(1, launches)
    achieved = r.bytes_per_frame * frames_per_launch / (kernel_ms / 1e3) / 1e9

    if args.profile:
        if rank == 0:
            real_stdout.write(json.dumps({"profile_mode": True, "workload": name, "launches": int(launches),
                                          "ms_per_step": ms / steps}) + "\n")
            real_stdout.flush()
        return

    # the same K steps from a cool start (W warm-ups only): what a short burst reaches at full
    # clock — the way MEASURED_PEAKS.json's copy peak itself was taken (best of 10 short copies)
    time.sleep(0.5)
    b_ms, _, b_steps = time_device(r, args.steps, args.warmup, use_dist, soak_s=0.0)
    burst = {"value": r.batch * b_steps * world / (b_ms / 1e3), "unit": "frames/s",
             "frac_of_hbm_peak": r.bytes_per_frame * r.batch * b_steps / (b_ms / 1e3) / 1e9 / len(devices) / peak,
             "note": "same K steps after 0.5 s idle + W warm-ups only (boost clocks, not power-capped)"}

    # e2e: pinned host frames through the same public call
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    r.prepare_host(max(1, args.e2e_batch))
    e_ms, st = time_host(r, e2e_steps, args.warmup, use_dist)
    e2e_value = r.e2e_batch * e2e_steps * world / (e_ms / 1e3)
    q_ms = time_host_queued(r, e2e_steps, args.warmup, use_dist)
    pcie_peak = pcie_bidir_peak_gbs()

    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))
        ent = tj.get(f"{name}/{args.content}") or tj.get(name)
        if isinstance(ent, dict):
            # captured per 16-frame launch; scaled to this run's frames per launch
            traffic = ent["bytes_per_frame"] * frames_per_launch
            traffic_src = ent["source"]

    spec = WORKLOADS[name]
    line = {
        "metric": metric_name(name),
        "value": value, "unit": "frames/s", "n_gpus": n_gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": ms / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(name, args.content),
        "run": {"frames_per_step": r.batch, "launches_per_step": launches / steps, "timed_ms": ms,
                "l2_hygiene": "inputs larger than L2 (batch in+out = %d MB per GPU)" %
                              (r.batch // len(devices) * r.bytes_per_frame // 1000000),
                "parallelism": (f"frame-parallel x{n_gpus}, no collective, " +
                                ("one process, b200vf_group (one host thread + 3 streams per device)"
                                 if args.single_process else "one process per GPU")),
                "kernel": kernel_note,
                **({"in_place_inputs": "every timed step filters buffers holding pristine content"}
                   if r.inplace else {})},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": r.bytes_per_frame * frames_per_launch,
                     "algorithmic_bytes_per_pixel": spec["bytes_per_pixel"],
                     "kernel_ms": kernel_ms,
                     "timed_after": "W warm-up steps + 0.3 s of the same kernel (steady-state clocks)"},
        "burst": burst,
        "e2e": {"value": e2e_value, "unit": "frames/s",
                "h2d_bytes_per_step": st["h2d_bytes"] // e2e_steps,
                "d2h_bytes_per_step": st["d2h_bytes"] // e2e_steps,
                "frames_per_step": r.e2e_batch, "steps": e2e_steps,
                "pcie_gbs_each_way_per_gpu": st["h2d_bytes"] / (e_ms / 1e3) / 1e9 / len(devices),
                "pcie_bidir_peak_gbs_each_way": pcie_peak,
                "frac_of_pcie_peak": st["h2d_bytes"] / (e_ms / 1e3) / 1e9 / len(devices) / pcie_peak,
                "calls": "synchronous: every call returns with its frames complete (the reference's transform_frame)",
                "queued": {"value": r.e2e_batch * e2e_steps * world / (q_ms / 1e3), "unit": "frames/s",
                           "calls": "\"host.async\" = 1, one call of latency: call k is queued, then call "
                                    "k-1's ticket is waited for (an element that holds one buffer back)",
                           "frac_of_pcie_peak": st["h2d_bytes"] / (q_ms / 1e3) / 1e9 / len(devices) / pcie_peak}},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    r.close()
    del r
    torch.cuda.empty_cache()

    # the headline workload on every content class, worst class named beside the headline
    classes = {args.content: {k: res[k] for k in ("frames_per_s", "gbs_per_gpu", "frac_of_hbm_peak", "kernel")}}
    if not args.no_classes:
        for content in (spec["contents"] or ()):
            if content in classes:
                continue
            cres, cr, *_ = measure_workload(g, eng, name, content, min(batch, 64), max(5, args.steps), 3,
                                            use_dist, rank, world, peak)
            classes[content] = {k: cres[k] for k in ("frames_per_s", "gbs_per_gpu", "frac_of_hbm_peak", "kernel")}
            cr.close()
            del cr
            torch.cuda.empty_cache()
    worst = min(classes, key=lambda c: classes[c]["frac_of_hbm_peak"])
    line["content_classes"] = classes
    line["worst_class"] = {"content": worst, **classes[worst],
                           "note": "min over the content classes of SURVEY.md §8(d) + noise (grad +-2 codes)"}

    if not args.no_extras and not args.single_process:
        extras = {}
        for wn, ws in WORKLOADS.items():
            for content in (ws["contents"] or (args.content,)):
                if wn == name:
                    continue
                b = default_batch(wn)
                k = max(5, args.steps)
                eres, er, *_ = measure_workload(g, eng, wn, content, b, k, 3, use_dist, rank, world, peak)
                extras[f"{wn}/{content}"] = eres
                if wn == "hsvdetector_4k":  # cfg4: system-memory frames through the pipeline
                    er.prepare_host(8)
                    hms, hst = time_host(er, 5, 1, use_dist)
                    extras[f"{wn}/{content}"]["e2e_frames_per_s"] = 8 * 5 * world / (hms / 1e3)
                    extras[f"{wn}/{content}"]["e2e_pcie_gbs_each_way"] = \
                        hst["h2d_bytes"] * world / (hms / 1e3) / 1e9
                er.close()
                del er
                torch.cuda.empty_cache()
        line["workloads"] = extras

    if rank == 0 and n_gpus == 1:
        line["cpu_baseline"] = cpu_baseline(name, args.content)
    if rank == 0:
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    if use_dist:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------
# CPU arm: the reference restated (oracle/) on the host cores
# ----------------------------------------------------------------------------------------
def cpu_run(name, content, n_frames, n_threads):
    """Process n_frames of the workload frame-parallel on n_threads; returns seconds."""
    import oracle
    s = WORKLOADS[name]
    elem, w, h = s["element"], s["width"], s["height"]
    lut = oracle.Lut(text=lut_text(s)) if s["lut"] else None
    wide, bpp3 = s["in_fmt"].startswith("RGBA64"), s["in_fmt"] in ("RGB", "BGR")
    uniq = [workload_frame(content, w, h, i, wide, bpp3) for i in range(min(n_frames, 4))]
    srcs = [uniq[i % len(uniq)].copy() for i in range(n_frames)]
    out_bytes = w * h * {"RGB": 3, "BGR": 3, "RGBA64_LE": 8, "RGBA64_BE": 8}.get(s["out_fmt"], 4)
    dsts = [np.empty(out_bytes, np.uint8) for _ in srcs]
    t0 = time.perf_counter()
    if elem == "colorlut_convert":  # the colorlut step alone, on the same bytes taken as RGBA
        rc = oracle.colorlut_frames_mt(lut, srcs, dsts, w, h, "RGBA", n_threads)
    elif elem == "colorlut":
        rc = oracle.colorlut_frames_mt(lut, srcs, dsts, w, h, s["in_fmt"], n_threads)
    elif elem == "hsvfilter":
        rc = oracle.hsvfilter_frames_mt(srcs, w, h, s["in_fmt"], CFG2, n_threads)
    elif elem == "hsvdetector":
        rc = oracle.hsvdetector_frames_mt(srcs, dsts, w, h, s["in_fmt"], s["out_fmt"], DET_CFG4, n_threads)
    else:  # chain / pipelines = the two elements back to back, as the reference pipeline runs them
        rc = oracle.colorlut_frames_mt(lut, srcs, dsts, w, h, "RGBA", n_threads)
        rc |= oracle.hsvfilter_frames_mt(dsts, w, h, "RGBA", CFG2, n_threads)
    dt = time.perf_counter() - t0
    assert rc == 0
    return dt


def cpu_baseline(name, content):
    """Bounded sample (~10-30 s of CPU work) of the same workload on the box's host cores."""
    import oracle
    oracle.build()
    cores = os.cpu_count() or 1
    t1 = cpu_run(name, content, 2, 1)                       # faithful: one streaming thread
    n = cores * 2
    tn = cpu_run(name, content, n, cores)                   # generous: one pipeline per core
    s = WORKLOADS[name]
    return {"value": n / tn, "unit": "frames/s", "cores": cores, "kind": "port",
            "sample": f"{n} frames of {name}/{content} frame-parallel on {cores} threads "
                      f"(plus 2 frames on 1 thread)",
            "value_1thread": 2 / t1, "ns_per_pixel_1thread": t1 / 2 / (s["width"] * s["height"]) * 1e9}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    oracle.build()
    cores = os.cpu_count() or 1
    name = args.workload
    per_step = cores  # one frame per host thread per step: a bounded sample of the B200 arm's batch
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_run(name, args.content, per_step, cores)
    steps = max(1, args.steps)
    budget_s = 150.0
    t_total, done = 0.0, 0
    for _ in range(steps):
        t_total += cpu_run(name, args.content, per_step, cores)
        done += 1
        if t_total > budget_s:
            break
    value = per_step * done / t_total
    line = {
        "impl": "reference",
        "metric": metric_name(name),
        "value": value, "unit": "frames/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
        "steps": done, "warmup": args.warmup, "ms_per_step": t_total / done * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": workload_config(name, args.content),
        "run": {"frames_per_step": per_step,
                "note": "CPU restatement of the reference (oracle/); Rust toolchain absent; each step "
                        "is a bounded sample of the workload (one frame per host thread)"},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": f"{per_step} frames/step x {done} steps, frame-parallel on "
                                   f"{cores} threads"},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=HEADLINE, choices=sorted(WORKLOADS))
    ap.add_argument("--content", default="grad", choices=list(CONTENT_CLASSES))
    ap.add_argument("--batch", type=int, default=0, help="frames per step and GPU (0 = per workload)")
    ap.add_argument("--e2e-batch", type=int, default=8)
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-extras", action="store_true", help="skip the other workloads")
    ap.add_argument("--no-classes", action="store_true", help="skip the other content classes")
    ap.add_argument("--single-process", action="store_true",
                    help="drive --gpus N devices from this one process through b200vf_group")
    ap.add_argument("--hsv-path", type=int, default=0, choices=[0, 1, 2],
                    help="\"hsv.path\" option for the HSV / chain workloads (0 = auto; under a profiler "
                         "the auto policy's own timings are meaningless, so captures pin 1 or 2)")
    ap.add_argument("--option", action="append", default=[], metavar="KEY=VALUE",
                    help="library option on top of the workload's, e.g. lut.path=4 to pin a kernel under ncu")
    ap.add_argument("--profile", action="store_true",
                    help="for ncu: exactly W + K steps of the workload's kernel, nothing else")
    args = ap.parse_args()
    global PROFILE_MODE, HSV_PATH
    PROFILE_MODE = args.profile
    HSV_PATH = args.hsv_path
    for kv in args.option:
        k, _, v = kv.partition("=")
        EXTRA_OPTIONS[k] = int(v)
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
