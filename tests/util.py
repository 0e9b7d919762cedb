"""Helpers for the parity tests: move numpy frames through the C ABI."""
import numpy as np

import gst_plugins_rs_b200 as g
from gst_plugins_rs_b200.api import BYTES_PER_PIXEL, frame_of

CFG2 = (37.5, 1.2, 0.05, 0.9, 0.02)          # SURVEY.md §8(d) cfg2 hsvfilter settings
IDENTITY = (0.0, 1.0, 0.0, 1.0, 0.0)         # hsvfilter/imp.rs:25-29 defaults
DET_DEFAULT = (0.0, 10.0, 0.0, 0.15, 0.0, 0.3)   # hsvdetector/imp.rs:26-31 defaults
DET_CFG4 = (120.0, 30.0, 0.6, 0.4, 0.6, 0.4)


def _buffers(arr, memory):
    """Return (holder, frame-able object) for a flat uint8 numpy array in the requested memory."""
    import torch
    flat = np.ascontiguousarray(arr, np.uint8).reshape(-1)
    if memory == "device":
        t = torch.from_numpy(flat.copy()).cuda()
        return t
    if memory == "pinned":
        t = torch.from_numpy(flat.copy()).pin_memory()
        return t
    if memory == "host":
        return flat.copy()
    raise ValueError(memory)


def _to_numpy(buf):
    if isinstance(buf, np.ndarray):
        return buf
    return buf.cpu().numpy()


def gpu_hsvfilter(ctx, arr, width, height, fmt, params, stride=None, memory="device"):
    buf = _buffers(arr, memory)
    f = frame_of(buf, width, height, fmt, stride)
    ctx.hsvfilter(f, g.HsvFilterParams(*params))
    ctx.synchronize()
    return _to_numpy(buf).reshape(-1)


def gpu_hsvdetector(ctx, arr, width, height, in_fmt, out_fmt, params, in_stride=None,
                    out_stride=None, memory="device", fill=0xA5):
    out_stride = out_stride or width * 4
    src = _buffers(arr, memory)
    dst = _buffers(np.full(height * out_stride, fill, np.uint8), memory)
    ctx.hsvdetector(frame_of(src, width, height, in_fmt, in_stride),
                    frame_of(dst, width, height, out_fmt, out_stride),
                    g.HsvDetectorParams(*params))
    ctx.synchronize()
    return _to_numpy(dst).reshape(-1)


def gpu_colorlut(ctx, arr, width, height, fmt="RGBA", src_stride=None, dst_stride=None,
                 memory="device", fill=0xA5):
    bpp = BYTES_PER_PIXEL[fmt]
    src_stride = src_stride or width * bpp
    dst_stride = dst_stride or src_stride
    src = _buffers(arr, memory)
    dst = _buffers(np.full(height * dst_stride, fill, np.uint8), memory)
    ctx.colorlut(frame_of(src, width, height, fmt, src_stride),
                 frame_of(dst, width, height, fmt, dst_stride))
    ctx.synchronize()
    return _to_numpy(dst).reshape(-1)


def diff_report(a, b):
    """(max abs diff, exact-match fraction) of two uint8 arrays."""
    a = np.asarray(a).astype(np.int16)
    b = np.asarray(b).astype(np.int16)
    d = np.abs(a - b)
    return int(d.max()) if d.size else 0, float((d == 0).mean()) if d.size else 1.0
