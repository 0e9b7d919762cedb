"""Timing of the non-headline formats/paths (dev aid): RGB/BGR 3-byte, RGBA64, 1D LUT, strided."""
import sys

import torch

sys.path.insert(0, ".")
import gst_plugins_rs_b200 as g
from gst_plugins_rs_b200 import frames
from gst_plugins_rs_b200.api import frame_array, frame_of

PEAK = 6548.5


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    w, h, nb = 3840, 2160, 8
    ctx = g.Context(0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    rows = []

    def bufs(bpp, stride=None):
        stride = stride or w * bpp
        return [torch.from_numpy(frames.random_bytes(stride * h, i)).cuda() for i in range(nb)], stride

    p = g.HsvFilterParams(37.5, 1.2, 0.05, 0.9, 0.02)
    dp = g.HsvDetectorParams(120, 30, 0.6, 0.4, 0.6, 0.4)
    b3, s3 = bufs(3)
    f3 = frame_array([frame_of(t, w, h, "RGB", s3) for t in b3])
    rows.append(("hsvfilter RGB (6 B/px)", timed(lambda: ctx.hsvfilter_batch(f3, p)), 6))
    o4, _ = bufs(4)
    fo4 = frame_array([frame_of(t, w, h, "RGBA") for t in o4])
    rows.append(("hsvdetector RGB->RGBA (7 B/px)", timed(lambda: ctx.hsvdetector_batch(f3, fo4, dp)), 7))
    b4, s4 = bufs(4, w * 4 + 64)
    f4 = frame_array([frame_of(t, w, h, "RGBA", s4) for t in b4])
    rows.append(("hsvfilter RGBA stride+64 (8 B/px)", timed(lambda: ctx.hsvfilter_batch(f4, p)), 8))
    b4u, s4u = bufs(4, w * 4 + 4)
    f4u = frame_array([frame_of(t, w, h, "RGBA", s4u) for t in b4u])
    rows.append(("hsvfilter RGBA stride+4 unaligned (8 B/px)", timed(lambda: ctx.hsvfilter_batch(f4u, p)), 8))
    b8, s8 = bufs(8)
    o8 = [torch.empty_like(t) for t in b8]
    ctx.set_lut_from_cube(g.parse_cube(frames.cube_text_3d(33)))
    for fmt in ("RGBA64_LE", "RGBA64_BE"):
        fi = frame_array([frame_of(t, w, h, fmt) for t in b8])
        fo = frame_array([frame_of(t, w, h, fmt) for t in o8])
        rows.append((f"colorlut33 {fmt} rand (16 B/px)", timed(lambda: ctx.colorlut_batch(fi, fo)), 16))
    g16 = frames.frame_grad(w, h).reshape(-1, 4).astype("<u2") * 257
    t16 = [torch.from_numpy(g16.view("u1").reshape(-1).copy()).cuda() for _ in range(nb)]
    fi = frame_array([frame_of(t, w, h, "RGBA64_LE") for t in t16])
    fo = frame_array([frame_of(t, w, h, "RGBA64_LE") for t in o8])
    rows.append(("colorlut33 RGBA64_LE grad (16 B/px)", timed(lambda: ctx.colorlut_batch(fi, fo)), 16))
    ctx.set_lut_from_cube(g.parse_cube(frames.cube_text_1d(1024)))
    gr = [torch.from_numpy(frames.frame_grad(w, h).reshape(-1).copy()).cuda() for _ in range(nb)]
    fi = frame_array([frame_of(t, w, h, "RGBA") for t in gr])
    rows.append(("colorlut 1D-1024 RGBA grad (8 B/px)", timed(lambda: ctx.colorlut_batch(fi, fo4)), 8))
    rows.append(("colorlut 1D-1024 RGBA64 grad (16 B/px)",
                 timed(lambda: ctx.colorlut_batch(frame_array([frame_of(t, w, h, "RGBA64_LE") for t in t16]), fo)), 16))
    for name, ms, bpp in rows:
        gbs = bpp * w * h * nb / ms / 1e6
        print(f"{name:46s} {ms:8.3f} ms  {gbs:8.1f} GB/s  {gbs / PEAK * 100:5.1f}%  {nb / ms * 1e3:8.0f} f/s")


if __name__ == "__main__":
    main()
