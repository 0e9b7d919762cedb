"""ctypes loader for libb200vf.so (the C ABI declared in include/b200vf.h).

The library is the product: if it is missing or fails to load we raise — there is
no Python / CPU fallback anywhere in this package.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200VF_LIB") or os.path.join(_HERE, "libb200vf.so")  # override: A/B builds


class Frame(C.Structure):
    """b200vf_frame"""
    _fields_ = [("data", C.c_void_p), ("stride", C.c_int64), ("width", C.c_uint32),
                ("height", C.c_uint32), ("format", C.c_uint32), ("memory", C.c_uint32)]


class HsvFilterParams(C.Structure):
    """b200vf_hsvfilter_params — defaults hsvfilter/imp.rs:25-29"""
    _fields_ = [("hue_shift", C.c_float), ("saturation_mul", C.c_float),
                ("saturation_off", C.c_float), ("value_mul", C.c_float), ("value_off", C.c_float)]


class HsvDetectorParams(C.Structure):
    """b200vf_hsvdetector_params — defaults hsvdetector/imp.rs:26-31"""
    _fields_ = [("hue_ref", C.c_float), ("hue_var", C.c_float), ("saturation_ref", C.c_float),
                ("saturation_var", C.c_float), ("value_ref", C.c_float), ("value_var", C.c_float)]


class Cube(C.Structure):
    """b200vf_cube"""
    _fields_ = [("kind", C.c_uint32), ("size", C.c_uint32), ("domain_scale", C.c_float * 3),
                ("domain_offset", C.c_float * 3), ("data", C.POINTER(C.c_float)),
                ("n_floats", C.c_size_t)]


class Stats(C.Structure):
    """b200vf_stats"""
    _fields_ = [("kernel_launches", C.c_uint64), ("frames", C.c_uint64),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64)]


class PoolConfig(C.Structure):
    """b200vf_pool_config"""
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("format", C.c_uint32),
                ("min_buffers", C.c_uint32), ("max_buffers", C.c_uint32),
                ("host_pinned", C.c_uint32)]


class PoolStats(C.Structure):
    """b200vf_pool_stats"""
    _fields_ = [("allocated", C.c_uint32), ("outstanding", C.c_uint32),
                ("frame_bytes", C.c_uint64), ("stride", C.c_int64)]


POOL_DONTWAIT = 1

_P = C.POINTER
_ctx = C.c_void_p
_pool = C.c_void_p
_group = C.c_void_p

# name -> (restype, argtypes).  Must list every symbol include/b200vf.h declares;
# tests/test_abi_surface.py checks header, this table and the .so against each other.
PROTOTYPES = {
    "b200vf_version": (C.c_char_p, []),
    "b200vf_status_string": (C.c_char_p, [C.c_int]),
    "b200vf_device_count": (C.c_int, [_P(C.c_int)]),
    "b200vf_format_bytes_per_pixel": (C.c_uint32, [C.c_uint32]),
    "b200vf_format_name": (C.c_char_p, [C.c_uint32]),
    "b200vf_format_from_name": (C.c_int, [C.c_char_p]),
    "b200vf_ctx_create": (C.c_int, [C.c_int, _P(_ctx)]),
    "b200vf_ctx_destroy": (None, [_ctx]),
    "b200vf_last_error": (C.c_char_p, [_ctx]),
    "b200vf_ctx_device": (C.c_int, [_ctx]),
    "b200vf_ctx_synchronize": (C.c_int, [_ctx]),
    "b200vf_ctx_get_stream": (C.c_void_p, [_ctx]),
    "b200vf_ctx_set_stream": (C.c_int, [_ctx, C.c_void_p]),
    "b200vf_ctx_wait_for": (C.c_int, [_ctx, _ctx]),
    "b200vf_ctx_host_ticket": (C.c_uint64, [_ctx]),
    "b200vf_ctx_host_wait": (C.c_int, [_ctx, C.c_uint64]),
    "b200vf_ctx_set_option": (C.c_int, [_ctx, C.c_char_p, C.c_int64]),
    "b200vf_ctx_get_option": (C.c_int, [_ctx, C.c_char_p, _P(C.c_int64)]),
    "b200vf_ctx_get_stats": (C.c_int, [_ctx, _P(Stats)]),
    "b200vf_ctx_reset_stats": (C.c_int, [_ctx]),
    "b200vf_host_alloc": (C.c_int, [C.c_size_t, _P(C.c_void_p)]),
    "b200vf_host_free": (C.c_int, [C.c_void_p]),
    "b200vf_host_is_pinned": (C.c_int, [C.c_void_p]),
    "b200vf_ctx_host_memory_released": (C.c_int, [_ctx, C.c_void_p, C.c_size_t]),
    "b200vf_device_alloc": (C.c_int, [_ctx, C.c_size_t, _P(C.c_void_p)]),
    "b200vf_device_free": (C.c_int, [_ctx, C.c_void_p]),
    "b200vf_memcpy": (C.c_int, [_ctx, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]),
    "b200vf_cube_parse": (C.c_int, [C.c_char_p, C.c_size_t, _P(Cube), C.c_char_p, C.c_size_t]),
    "b200vf_cube_parse_file": (C.c_int, [C.c_char_p, _P(Cube), C.c_char_p, C.c_size_t]),
    "b200vf_cube_free": (None, [_P(Cube)]),
    "b200vf_colorlut_set_lut": (C.c_int, [_ctx, C.c_uint32, C.c_uint32, _P(C.c_float),
                                          _P(C.c_float), _P(C.c_float)]),
    "b200vf_colorlut_set_lut_file": (C.c_int, [_ctx, C.c_char_p]),
    "b200vf_colorlut_clear_lut": (C.c_int, [_ctx]),
    "b200vf_colorlut_process": (C.c_int, [_ctx, _P(Frame), _P(Frame)]),
    "b200vf_colorlut_process_batch": (C.c_int, [_ctx, _P(Frame), _P(Frame), C.c_size_t]),
    "b200vf_colorlut_convert_process_batch": (C.c_int, [_ctx, _P(Frame), _P(Frame), C.c_size_t]),
    "b200vf_hsvfilter_process": (C.c_int, [_ctx, _P(Frame), _P(HsvFilterParams)]),
    "b200vf_hsvfilter_process_batch": (C.c_int, [_ctx, _P(Frame), C.c_size_t,
                                                 _P(HsvFilterParams)]),
    "b200vf_hsvdetector_process": (C.c_int, [_ctx, _P(Frame), _P(Frame), _P(HsvDetectorParams)]),
    "b200vf_hsvdetector_process_batch": (C.c_int, [_ctx, _P(Frame), _P(Frame), C.c_size_t,
                                                   _P(HsvDetectorParams)]),
    "b200vf_group_create": (C.c_int, [_P(C.c_int), C.c_size_t, _P(_group)]),
    "b200vf_group_destroy": (None, [_group]),
    "b200vf_group_size": (C.c_size_t, [_group]),
    "b200vf_group_ctx": (_ctx, [_group, C.c_size_t]),
    "b200vf_group_last_error": (C.c_char_p, [_group]),
    "b200vf_group_set_option": (C.c_int, [_group, C.c_char_p, C.c_int64]),
    "b200vf_group_synchronize": (C.c_int, [_group]),
    "b200vf_group_colorlut_set_lut": (C.c_int, [_group, C.c_uint32, C.c_uint32, _P(C.c_float),
                                                _P(C.c_float), _P(C.c_float)]),
    "b200vf_group_colorlut_set_lut_file": (C.c_int, [_group, C.c_char_p]),
    "b200vf_group_colorlut_clear_lut": (C.c_int, [_group]),
    "b200vf_group_colorlut_process_batch": (C.c_int, [_group, _P(Frame), _P(Frame), C.c_size_t]),
    "b200vf_group_hsvfilter_process_batch": (C.c_int, [_group, _P(Frame), C.c_size_t,
                                                       _P(HsvFilterParams)]),
    "b200vf_group_hsvdetector_process_batch": (C.c_int, [_group, _P(Frame), _P(Frame), C.c_size_t,
                                                         _P(HsvDetectorParams)]),
    "b200vf_group_chain_lut_hsv_process_batch": (C.c_int, [_group, _P(Frame), _P(Frame), C.c_size_t,
                                                           _P(HsvFilterParams)]),
    "b200vf_pool_create": (C.c_int, [C.c_int, _P(PoolConfig), _P(_pool)]),
    "b200vf_pool_destroy": (None, [_pool]),
    "b200vf_pool_acquire": (C.c_int, [_pool, C.c_uint32, _P(Frame)]),
    "b200vf_pool_release": (C.c_int, [_pool, _P(Frame), C.c_void_p]),
    "b200vf_pool_release_after": (C.c_int, [_pool, _P(Frame), _ctx]),
    "b200vf_pool_get_stats": (C.c_int, [_pool, _P(PoolStats)]),
    "b200vf_pool_device": (C.c_int, [_pool]),
    "b200vf_pointer_info": (C.c_int, [C.c_void_p, _P(C.c_uint32), _P(C.c_int)]),
    "b200vf_debug_table_indices": (C.c_int, [_P(C.c_uint32), C.c_size_t, _P(C.c_uint32)]),
    "b200vf_debug_hsv_from_rgb": (C.c_int, [_ctx, C.c_void_p, C.c_size_t, C.c_void_p]),
    "b200vf_chain_lut_hsv_process_batch": (C.c_int, [_ctx, _P(Frame), _P(Frame), C.c_size_t,
                                                     _P(HsvFilterParams)]),
}

_lib = None


def load():
    """Load libb200vf.so and bind every prototype.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (or make -C gst-plugins-rs_b200/csrc).  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError = ABI drift, fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
