"""Thin Python view of the C ABI (include/b200vf.h) for tests and bench.py.

Nothing here computes pixels: every call goes straight into libb200vf.so.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import Cube, Frame, HsvDetectorParams, HsvFilterParams, Stats

# b200vf_format
FORMATS = {"RGBA": 0, "RGBx": 1, "xRGB": 2, "ARGB": 3, "BGRx": 4, "BGRA": 5, "xBGR": 6, "ABGR": 7,
           "RGB": 8, "BGR": 9, "RGBA64_LE": 10, "RGBA64_BE": 11}
BYTES_PER_PIXEL = {"RGBA": 4, "RGBx": 4, "xRGB": 4, "ARGB": 4, "BGRx": 4, "BGRA": 4, "xBGR": 4,
                   "ABGR": 4, "RGB": 3, "BGR": 3, "RGBA64_LE": 8, "RGBA64_BE": 8}
MEM_HOST, MEM_DEVICE = 0, 1
LUT_1D, LUT_3D = 1, 3

OK = 0
ERR_INVALID_ARG, ERR_UNSUPPORTED_FORMAT, ERR_CUDA, ERR_NO_LUT = -1, -2, -3, -4
ERR_PARSE, ERR_IO, ERR_NO_DEVICE, ERR_NOMEM, ERR_SETTINGS = -5, -6, -7, -8, -9


class B200VFError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"b200vf status {status}: {message}")
        self.status = status
        self.message = message


def frame_of(buf, width, height, fmt, stride=None, offset=0):
    """Describe `buf` (torch tensor, numpy array or raw int pointer) as a b200vf_frame."""
    bpp = BYTES_PER_PIXEL[fmt]
    if stride is None:
        stride = width * bpp
    if isinstance(buf, np.ndarray):
        ptr, mem = buf.ctypes.data, MEM_HOST
    elif isinstance(buf, int):
        ptr, mem = buf, MEM_DEVICE
    else:  # torch tensor
        ptr = buf.data_ptr()
        mem = MEM_DEVICE if buf.is_cuda else MEM_HOST
    return Frame(ptr + offset, stride, width, height, FORMATS[fmt], mem)


def parse_cube(text):
    """b200vf_cube_parse → dict(kind,size,scale,offset,data) or raises B200VFError."""
    lib = _lib.load()
    raw = text.encode() if isinstance(text, str) else bytes(text)
    cube, err = Cube(), C.create_string_buffer(512)
    rc = lib.b200vf_cube_parse(raw, len(raw), C.byref(cube), err, len(err))
    if rc != OK:
        raise B200VFError(rc, err.value.decode(errors="replace"))
    try:
        return _cube_to_dict(cube)
    finally:
        lib.b200vf_cube_free(C.byref(cube))


def parse_cube_file(path):
    lib = _lib.load()
    cube, err = Cube(), C.create_string_buffer(512)
    rc = lib.b200vf_cube_parse_file(str(path).encode(), C.byref(cube), err, len(err))
    if rc != OK:
        raise B200VFError(rc, err.value.decode(errors="replace"))
    try:
        return _cube_to_dict(cube)
    finally:
        lib.b200vf_cube_free(C.byref(cube))


def _cube_to_dict(cube):
    data = np.ctypeslib.as_array(cube.data, shape=(cube.n_floats,)).copy()
    return {"kind": int(cube.kind), "size": int(cube.size),
            "scale": np.array(list(cube.domain_scale), np.float32),
            "offset": np.array(list(cube.domain_offset), np.float32), "data": data}


class Context:
    """b200vf_ctx — one per element instance, bound to one device."""

    def __init__(self, device=0):
        self.lib = _lib.load()
        h = C.c_void_p()
        rc = self.lib.b200vf_ctx_create(device, C.byref(h))
        if rc != OK:
            raise B200VFError(rc, (self.lib.b200vf_last_error(None) or b"").decode())
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.b200vf_ctx_destroy(self.h)
            self.h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc != OK:
            raise B200VFError(rc, (self.lib.b200vf_last_error(self.h) or b"").decode())

    def synchronize(self):
        self._check(self.lib.b200vf_ctx_synchronize(self.h))

    def set_stream(self, cuda_stream):
        self._check(self.lib.b200vf_ctx_set_stream(self.h, C.c_void_p(cuda_stream)))

    def get_stream(self):
        return self.lib.b200vf_ctx_get_stream(self.h)

    def wait_for(self, upstream):
        """Stream-order this context after everything enqueued on `upstream` so far."""
        self._check(self.lib.b200vf_ctx_wait_for(self.h, upstream.h))

    def host_ticket(self):
        """Ticket of the most recent host-frame call (see the "host.async" option)."""
        return int(self.lib.b200vf_ctx_host_ticket(self.h))

    def host_wait(self, ticket):
        """Block until the host-frame call with this ticket and all earlier ones are complete."""
        self._check(self.lib.b200vf_ctx_host_wait(self.h, int(ticket)))

    def set_option(self, key, value):
        self._check(self.lib.b200vf_ctx_set_option(self.h, key.encode(), int(value)))

    def get_option(self, key):
        v = C.c_int64()
        self._check(self.lib.b200vf_ctx_get_option(self.h, key.encode(), C.byref(v)))
        return v.value

    def stats(self):
        s = Stats()
        self._check(self.lib.b200vf_ctx_get_stats(self.h, C.byref(s)))
        return {"kernel_launches": s.kernel_launches, "frames": s.frames,
                "h2d_bytes": s.h2d_bytes, "d2h_bytes": s.d2h_bytes}

    def reset_stats(self):
        self._check(self.lib.b200vf_ctx_reset_stats(self.h))

    def host_memory_released(self, ptr, nbytes=0):
        """b200vf_ctx_host_memory_released: call BEFORE freeing memory the context may have registered."""
        self._check(self.lib.b200vf_ctx_host_memory_released(self.h, C.c_void_p(ptr), nbytes))

    # ---- colorlut -----------------------------------------------------------
    def set_lut(self, kind, size, data, scale=(1, 1, 1), offset=(0, 0, 0)):
        data = np.ascontiguousarray(data, np.float32)
        sc = (C.c_float * 3)(*[float(x) for x in scale])
        of = (C.c_float * 3)(*[float(x) for x in offset])
        self._check(self.lib.b200vf_colorlut_set_lut(
            self.h, kind, size, data.ctypes.data_as(C.POINTER(C.c_float)), sc, of))

    def set_lut_from_cube(self, cube):
        self.set_lut(cube["kind"], cube["size"], cube["data"], cube["scale"], cube["offset"])

    def set_lut_file(self, location):
        loc = None if location is None else str(location).encode()
        self._check(self.lib.b200vf_colorlut_set_lut_file(self.h, loc))

    def clear_lut(self):
        self._check(self.lib.b200vf_colorlut_clear_lut(self.h))

    def colorlut(self, fin, fout):
        self._check(self.lib.b200vf_colorlut_process(self.h, C.byref(fin), C.byref(fout)))

    def colorlut_batch(self, fins, fouts):
        a, b = _arr(fins), _arr(fouts)
        self._check(self.lib.b200vf_colorlut_process_batch(self.h, a, b, len(fins)))

    def colorlut_convert_batch(self, fins, fouts):
        """colorlut with the videoconvert steps folded in: any 8-bit packed layout in, any out."""
        a, b = _arr(fins), _arr(fouts)
        self._check(self.lib.b200vf_colorlut_convert_process_batch(self.h, a, b, len(fins)))

    # ---- hsvfilter ----------------------------------------------------------
    def hsvfilter(self, frame, params):
        self._check(self.lib.b200vf_hsvfilter_process(self.h, C.byref(frame), C.byref(params)))

    def hsvfilter_batch(self, frames, params):
        a = _arr(frames)
        self._check(self.lib.b200vf_hsvfilter_process_batch(self.h, a, len(frames),
                                                            C.byref(params)))

    # ---- hsvdetector --------------------------------------------------------
    def hsvdetector(self, fin, fout, params):
        self._check(self.lib.b200vf_hsvdetector_process(self.h, C.byref(fin), C.byref(fout),
                                                        C.byref(params)))

    def hsvdetector_batch(self, fins, fouts, params):
        a, b = _arr(fins), _arr(fouts)
        self._check(self.lib.b200vf_hsvdetector_process_batch(self.h, a, b, len(fins),
                                                              C.byref(params)))

    # ---- colorlut ! hsvfilter ----------------------------------------------
    def chain_lut_hsv_batch(self, fins, fouts, params):
        a, b = _arr(fins), _arr(fouts)
        self._check(self.lib.b200vf_chain_lut_hsv_process_batch(self.h, a, b, len(fins),
                                                                C.byref(params)))


class _MemberView(Context):
    """A group member's context, borrowed (the group owns it)."""

    def __init__(self, lib, handle):
        self.lib, self.h = lib, handle

    def close(self):
        self.h = None

    __del__ = close


class Group:
    """b200vf_group — one process feeding several GPUs: frame i of a batch goes to member i mod G."""

    def __init__(self, devices):
        self.lib = _lib.load()
        devs = (C.c_int * len(devices))(*devices)
        self.h = C.c_void_p()
        rc = self.lib.b200vf_group_create(devs, len(devices), C.byref(self.h))
        if rc != OK:
            raise B200VFError(rc, (self.lib.b200vf_last_error(None) or b"").decode())
        self.devices = list(devices)

    def close(self):
        if getattr(self, "h", None):
            self.lib.b200vf_group_destroy(self.h)
            self.h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __len__(self):
        return self.lib.b200vf_group_size(self.h)

    def member(self, i):
        h = self.lib.b200vf_group_ctx(self.h, i)
        if not h:
            raise IndexError(i)
        return _MemberView(self.lib, C.c_void_p(h))

    def _check(self, rc):
        if rc != OK:
            raise B200VFError(rc, (self.lib.b200vf_group_last_error(self.h) or b"").decode())

    def set_option(self, key, value):
        self._check(self.lib.b200vf_group_set_option(self.h, key.encode(), int(value)))

    def synchronize(self):
        self._check(self.lib.b200vf_group_synchronize(self.h))

    def set_lut(self, kind, size, data, scale=(1, 1, 1), offset=(0, 0, 0)):
        data = np.ascontiguousarray(data, np.float32)
        sc = (C.c_float * 3)(*[float(x) for x in scale])
        of = (C.c_float * 3)(*[float(x) for x in offset])
        self._check(self.lib.b200vf_group_colorlut_set_lut(
            self.h, kind, size, data.ctypes.data_as(C.POINTER(C.c_float)), sc, of))

    def set_lut_from_cube(self, cube):
        self.set_lut(cube["kind"], cube["size"], cube["data"], cube["scale"], cube["offset"])

    def set_lut_file(self, location):
        loc = None if location is None else str(location).encode()
        self._check(self.lib.b200vf_group_colorlut_set_lut_file(self.h, loc))

    def clear_lut(self):
        self._check(self.lib.b200vf_group_colorlut_clear_lut(self.h))

    def colorlut_batch(self, fins, fouts):
        a, b = _arr(fins), _arr(fouts)
        self._check(self.lib.b200vf_group_colorlut_process_batch(self.h, a, b, len(fins)))

    def hsvfilter_batch(self, frames, params):
        a = _arr(frames)
        self._check(self.lib.b200vf_group_hsvfilter_process_batch(self.h, a, len(frames), C.byref(params)))

    def hsvdetector_batch(self, fins, fouts, params):
        a, b = _arr(fins), _arr(fouts)
        self._check(self.lib.b200vf_group_hsvdetector_process_batch(self.h, a, b, len(fins),
                                                                   C.byref(params)))

    def chain_lut_hsv_batch(self, fins, fouts, params):
        a, b = _arr(fins), _arr(fouts)
        self._check(self.lib.b200vf_group_chain_lut_hsv_process_batch(self.h, a, b, len(fins),
                                                                     C.byref(params)))


def debug_hsv_from_rgb(ctx, rgba_tensor, hsv_tensor):
    """Diagnostics: (h,s,v) floats of RGB→HSV for device RGBA pixels (see b200vf.h)."""
    n = rgba_tensor.numel() // 4
    ctx._check(ctx.lib.b200vf_debug_hsv_from_rgb(ctx.h, C.c_void_p(rgba_tensor.data_ptr()), n,
                                                 C.c_void_p(hsv_tensor.data_ptr())))


def _arr(frames):
    if isinstance(frames, C.Array):
        return frames
    return (Frame * len(frames))(*frames)


def frame_array(frames):
    """Pre-build the ctypes array once (bench hot loop)."""
    return _arr(list(frames))


class DevicePool:
    """b200vf_pool_*: pool of device frames of one geometry (memory:CUDAMemory buffer pool)."""

    def __init__(self, device, width, height, fmt, min_buffers=0, max_buffers=0, host_pinned=False):
        self.lib = _lib.load()
        self.h = C.c_void_p()
        cfg = _lib.PoolConfig(width, height, FORMATS[fmt] if isinstance(fmt, str) else fmt,
                              min_buffers, max_buffers, 1 if host_pinned else 0)
        rc = self.lib.b200vf_pool_create(device, C.byref(cfg), C.byref(self.h))
        if rc != OK:
            raise B200VFError(rc, (self.lib.b200vf_last_error(None) or b"").decode())

    def close(self):
        if self.h:
            self.lib.b200vf_pool_destroy(self.h)
            self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def acquire(self, dont_wait=False):
        f = Frame()
        rc = self.lib.b200vf_pool_acquire(self.h, _lib.POOL_DONTWAIT if dont_wait else 0, C.byref(f))
        if rc != OK:
            raise B200VFError(rc, (self.lib.b200vf_last_error(None) or b"").decode())
        return f

    def release(self, frame, last_use_stream=None):
        rc = self.lib.b200vf_pool_release(self.h, C.byref(frame), C.c_void_p(last_use_stream))
        if rc != OK:
            raise B200VFError(rc, (self.lib.b200vf_last_error(None) or b"").decode())

    def stats(self):
        st = _lib.PoolStats()
        self.lib.b200vf_pool_get_stats(self.h, C.byref(st))
        return {"allocated": st.allocated, "outstanding": st.outstanding,
                "frame_bytes": st.frame_bytes, "stride": st.stride}

    @property
    def device(self):
        return self.lib.b200vf_pool_device(self.h)


def host_is_pinned(ptr):
    return bool(_lib.load().b200vf_host_is_pinned(C.c_void_p(ptr)))


def pointer_info(ptr):
    """b200vf_pointer_info → (memory, device)."""
    lib = _lib.load()
    mem, dev = C.c_uint32(), C.c_int()
    rc = lib.b200vf_pointer_info(C.c_void_p(ptr), C.byref(mem), C.byref(dev))
    if rc != OK:
        raise B200VFError(rc, (lib.b200vf_last_error(None) or b"").decode())
    return mem.value, dev.value
