"""ctypes binding of oracle/liboracle.so (see vf_oracle.h).  Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")

FORMATS = {"RGBA": 0, "RGBx": 1, "xRGB": 2, "ARGB": 3, "BGRx": 4, "BGRA": 5, "xBGR": 6, "ABGR": 7,
           "RGB": 8, "BGR": 9, "RGBA64_LE": 10, "RGBA64_BE": 11}
BPP = {"RGBA": 4, "RGBx": 4, "xRGB": 4, "ARGB": 4, "BGRx": 4, "BGRA": 4, "xBGR": 4, "ABGR": 4,
       "RGB": 3, "BGR": 3, "RGBA64_LE": 8, "RGBA64_BE": 8}


class OrcCube(C.Structure):
    _fields_ = [("kind", C.c_int), ("size", C.c_uint32), ("domain_scale", C.c_float * 3),
                ("domain_offset", C.c_float * 3), ("data", C.POINTER(C.c_float)),
                ("n_floats", C.c_size_t)]


class FilterParams(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("hue_shift", "saturation_mul", "saturation_off",
                                         "value_mul", "value_off")]


class DetectorParams(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("hue_ref", "hue_var", "saturation_ref", "saturation_var",
                                         "value_ref", "value_var")]


def build(force=False):
    """Compile liboracle.so with the committed Makefile (gcc, -ffp-contract=off)."""
    if force or not os.path.exists(LIB_PATH) or any(
            os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(LIB_PATH)
            for f in ("vf_oracle.c", "vf_oracle.h", "Makefile")):
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        vp, sz, u32, i = C.c_void_p, C.c_size_t, C.c_uint32, C.c_int
        L.orc_cube_parse.argtypes = [C.c_char_p, sz, C.POINTER(OrcCube), C.c_char_p, sz]
        L.orc_cube_parse_file.argtypes = [C.c_char_p, C.POINTER(OrcCube), C.c_char_p, sz]
        L.orc_cube_free.argtypes = [C.POINTER(OrcCube)]
        L.orc_cube_free.restype = None
        L.orc_colorlut_frame.argtypes = [C.POINTER(OrcCube), vp, sz, vp, sz, u32, u32, i]
        L.orc_colorlut_frame_ex.argtypes = [C.POINTER(OrcCube), vp, sz, vp, sz, u32, u32, i, i]
        L.orc_hsvfilter_frame.argtypes = [vp, sz, u32, u32, i, C.POINTER(FilterParams)]
        L.orc_hsvdetector_frame.argtypes = [vp, sz, i, vp, sz, i, u32, u32,
                                            C.POINTER(DetectorParams)]
        L.orc_colorlut_frames_mt.argtypes = [C.POINTER(OrcCube), C.POINTER(vp), C.POINTER(vp), sz,
                                             sz, u32, u32, i, i]
        L.orc_hsvfilter_frames_mt.argtypes = [C.POINTER(vp), sz, sz, u32, u32, i,
                                              C.POINTER(FilterParams), i]
        L.orc_hsvdetector_frames_mt.argtypes = [C.POINTER(vp), C.POINTER(vp), sz, sz, i, sz, i,
                                                u32, u32, C.POINTER(DetectorParams), i]
        for n in ("orc_hsv_from_rgb", "orc_hsv_from_bgr"):
            getattr(L, n).argtypes = [C.POINTER(C.c_uint8), C.POINTER(C.c_float)]
            getattr(L, n).restype = None
        for n in ("orc_hsv_to_rgb", "orc_hsv_to_bgr"):
            getattr(L, n).argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_uint8)]
            getattr(L, n).restype = None
        L.orc_hsv_from_rgba_batch.argtypes = [vp, sz, vp]
        L.orc_hsv_from_rgba_batch.restype = None
        L.orc_colorlut_apply_u8.argtypes = [C.POINTER(OrcCube), C.POINTER(C.c_uint8),
                                            C.POINTER(C.c_uint8)]
        L.orc_colorlut_apply_u8.restype = None
        L.orc_colorlut_apply_u16.argtypes = [C.POINTER(OrcCube), C.POINTER(C.c_uint16),
                                             C.POINTER(C.c_uint16)]
        L.orc_colorlut_apply_u16.restype = None
        _lib = L
    return _lib


class CubeError(ValueError):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code  # 1 InvalidLut, 2 Io


class Lut:
    """Parsed CubeLut held by the oracle (parser.rs:68-74)."""

    def __init__(self, text=None, path=None):
        self.c = OrcCube()
        err = C.create_string_buffer(512)
        if path is not None:
            rc = lib().orc_cube_parse_file(str(path).encode(), C.byref(self.c), err, len(err))
        else:
            raw = text.encode() if isinstance(text, str) else bytes(text)
            rc = lib().orc_cube_parse(raw, len(raw), C.byref(self.c), err, len(err))
        if rc != 0:
            raise CubeError(rc, err.value.decode(errors="replace"))

    kind = property(lambda s: s.c.kind)
    size = property(lambda s: s.c.size)
    scale = property(lambda s: np.array(list(s.c.domain_scale), np.float32))
    offset = property(lambda s: np.array(list(s.c.domain_offset), np.float32))

    @property
    def data(self):
        return np.ctypeslib.as_array(self.c.data, shape=(self.c.n_floats,)).copy()

    def __del__(self):
        try:
            lib().orc_cube_free(C.byref(self.c))
        except Exception:
            pass


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def from_rgb(rgb):
    o = (C.c_float * 3)()
    lib().orc_hsv_from_rgb((C.c_uint8 * 3)(*rgb), o)
    return [o[0], o[1], o[2]]


def from_bgr(bgr):
    o = (C.c_float * 3)()
    lib().orc_hsv_from_bgr((C.c_uint8 * 3)(*bgr), o)
    return [o[0], o[1], o[2]]


def from_rgba_batch(rgba):
    """(n,4) uint8 → (n,3) float32 of orc_hsv_from_rgb."""
    rgba = np.ascontiguousarray(rgba, np.uint8).reshape(-1, 4)
    out = np.empty((len(rgba), 3), np.float32)
    lib().orc_hsv_from_rgba_batch(_ptr(rgba), len(rgba), _ptr(out))
    return out


def to_rgb(hsv):
    o = (C.c_uint8 * 3)()
    lib().orc_hsv_to_rgb((C.c_float * 3)(*hsv), o)
    return [o[0], o[1], o[2]]


def to_bgr(hsv):
    o = (C.c_uint8 * 3)()
    lib().orc_hsv_to_bgr((C.c_float * 3)(*hsv), o)
    return [o[0], o[1], o[2]]


INTERPOLATIONS = {"trilinear": 0, "tetrahedral": 1, "nearest": 2}


def colorlut(lut, src, width, height, fmt="RGBA", src_stride=None, dst_stride=None, dst=None,
             interpolation="trilinear"):
    """ColorLut::transform_frame on a (rows, stride) uint8 array; returns the output array.
    interpolation != "trilinear" is the extension without a reference counterpart."""
    bpp = BPP[fmt]
    src = np.ascontiguousarray(src, np.uint8)
    src_stride = src_stride or width * bpp
    dst_stride = dst_stride or src_stride
    if dst is None:
        dst = np.zeros(height * dst_stride, np.uint8)
    if interpolation == "trilinear":
        rc = lib().orc_colorlut_frame(C.byref(lut.c), _ptr(src), src_stride, _ptr(dst), dst_stride,
                                      width, height, FORMATS[fmt])
    else:
        rc = lib().orc_colorlut_frame_ex(C.byref(lut.c), _ptr(src), src_stride, _ptr(dst),
                                         dst_stride, width, height, FORMATS[fmt],
                                         INTERPOLATIONS[interpolation])
    if rc:
        raise ValueError("oracle colorlut: bad format/stride")
    return dst


def hsvfilter(data, width, height, fmt, params, stride=None):
    """HsvFilter::transform_frame_ip — modifies and returns a copy of `data`."""
    out = np.array(data, np.uint8, copy=True).reshape(-1)
    stride = stride or width * BPP[fmt]
    p = params if isinstance(params, FilterParams) else FilterParams(*params)
    rc = lib().orc_hsvfilter_frame(_ptr(out), stride, width, height, FORMATS[fmt], C.byref(p))
    if rc:
        raise ValueError("oracle hsvfilter: bad format")
    return out


def hsvdetector(src, width, height, in_fmt, out_fmt, params, in_stride=None, out_stride=None,
                dst=None):
    src = np.ascontiguousarray(src, np.uint8).reshape(-1)
    in_stride = in_stride or width * BPP[in_fmt]
    out_stride = out_stride or width * 4
    if dst is None:
        dst = np.zeros(height * out_stride, np.uint8)
    p = params if isinstance(params, DetectorParams) else DetectorParams(*params)
    rc = lib().orc_hsvdetector_frame(_ptr(src), in_stride, FORMATS[in_fmt], _ptr(dst), out_stride,
                                     FORMATS[out_fmt], width, height, C.byref(p))
    if rc:
        raise ValueError("oracle hsvdetector: bad format pair")
    return dst


def _ptr_array(arrays):
    return (C.c_void_p * len(arrays))(*[a.ctypes.data for a in arrays])


def colorlut_frames_mt(lut, srcs, dsts, width, height, fmt, n_threads):
    stride = width * BPP[fmt]
    return lib().orc_colorlut_frames_mt(C.byref(lut.c), _ptr_array(srcs), _ptr_array(dsts),
                                        len(srcs), stride, width, height, FORMATS[fmt], n_threads)


def hsvfilter_frames_mt(frames, width, height, fmt, params, n_threads):
    p = params if isinstance(params, FilterParams) else FilterParams(*params)
    return lib().orc_hsvfilter_frames_mt(_ptr_array(frames), len(frames), width * BPP[fmt], width,
                                         height, FORMATS[fmt], C.byref(p), n_threads)


def hsvdetector_frames_mt(srcs, dsts, width, height, in_fmt, out_fmt, params, n_threads):
    p = params if isinstance(params, DetectorParams) else DetectorParams(*params)
    return lib().orc_hsvdetector_frames_mt(_ptr_array(srcs), _ptr_array(dsts), len(srcs),
                                           width * BPP[in_fmt], FORMATS[in_fmt], width * 4,
                                           FORMATS[out_fmt], width, height, C.byref(p), n_threads)
