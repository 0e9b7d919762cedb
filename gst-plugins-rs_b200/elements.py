"""Python handle on the C++ element layer (elements/vf_elements.hpp): drive `colorlut`,
`hsvfilter`, `hsvdetector` the way a GStreamer pipeline would — set properties, start,
push frames through transform_frame / transform_frame_ip, stop.  No pixel code here."""
import ctypes as C
import json
import os

from . import _lib
from ._lib import Frame

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200vf_elements.so")

FLOW_OK, FLOW_ERROR = 0, -5
RESOURCE_ERRORS = {0: None, 1: "Settings", 2: "Read", 3: "Failed"}

_elib = None


def load():
    global _elib
    if _elib is None:
        _lib.load()  # libb200vf.so first (fails loudly if missing)
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run __graft_entry__.build()")
        L = C.CDLL(LIB_PATH)
        p, s, i, f = C.c_void_p, C.c_char_p, C.c_int, C.c_float
        L.b200vf_element_new.restype, L.b200vf_element_new.argtypes = p, [s, i]
        L.b200vf_element_free.restype, L.b200vf_element_free.argtypes = None, [p]
        L.b200vf_element_set_float.argtypes = [p, s, f]
        L.b200vf_element_set_string.argtypes = [p, s, s]
        L.b200vf_element_get_float.argtypes = [p, s, C.POINTER(f)]
        L.b200vf_element_get_string.restype, L.b200vf_element_get_string.argtypes = s, [p, s]
        L.b200vf_element_start.argtypes = [p]
        L.b200vf_element_stop.argtypes = [p]
        L.b200vf_element_message.restype, L.b200vf_element_message.argtypes = s, [p]
        L.b200vf_element_transform_frame.argtypes = [p, C.POINTER(Frame), C.POINTER(Frame)]
        L.b200vf_element_transform_frame_ip.argtypes = [p, C.POINTER(Frame)]
        L.b200vf_element_set_frames_in_flight.argtypes = [p, C.c_uint]
        L.b200vf_element_submit_input_frame.argtypes = [p, C.POINTER(Frame), C.POINTER(Frame)]
        L.b200vf_element_generate_output.argtypes = [p, C.POINTER(Frame)]
        L.b200vf_element_drain.argtypes = [p, C.POINTER(Frame)]
        L.b200vf_element_transform_caps.restype = s
        L.b200vf_element_transform_caps.argtypes = [p, i, s, s]
        L.b200vf_element_describe.restype, L.b200vf_element_describe.argtypes = s, [s]
        L.b200vf_element_context.restype, L.b200vf_element_context.argtypes = p, [p]
        _elib = L
    return _elib


def describe(factory_name):
    """Element surface as a dict (same facts as docs/plugins/gst_plugins_cache.json)."""
    return json.loads(load().b200vf_element_describe(factory_name.encode()).decode())


class ElementError(RuntimeError):
    def __init__(self, domain, message):
        super().__init__(f"{domain}: {message}")
        self.domain = domain
        self.message = message


class Element:
    """gst::ElementFactory::make(name) + the calls GstBaseTransform / GstVideoFilter issue."""

    def __init__(self, factory_name, device=0, **props):
        self.L = load()
        self.h = self.L.b200vf_element_new(factory_name.encode(), device)
        if not self.h:
            raise ValueError(f"no such element factory: {factory_name}")
        self.name = factory_name
        for k, v in props.items():
            if not self.set_property(k.replace("_", "-"), v):
                raise ValueError(f"cannot set {k}={v!r}")

    def close(self):
        if getattr(self, "h", None):
            self.L.b200vf_element_free(self.h)
            self.h = None

    __del__ = close

    def set_property(self, name, value):
        if isinstance(value, (str, bytes, os.PathLike)) or value is None:
            v = None if value is None else os.fspath(value).encode()
            return bool(self.L.b200vf_element_set_string(self.h, name.encode(), v))
        return bool(self.L.b200vf_element_set_float(self.h, name.encode(), float(value)))

    def get_property(self, name):
        out = C.c_float()
        if self.L.b200vf_element_get_float(self.h, name.encode(), C.byref(out)):
            return out.value
        s = self.L.b200vf_element_get_string(self.h, name.encode())
        return None if s is None else s.decode()

    def start(self):
        rc = self.L.b200vf_element_start(self.h)
        if rc:
            raise ElementError(RESOURCE_ERRORS.get(rc, rc),
                               self.L.b200vf_element_message(self.h).decode())

    def stop(self):
        self.L.b200vf_element_stop(self.h)

    def message(self):
        return self.L.b200vf_element_message(self.h).decode()

    def transform_frame(self, fin, fout):
        return self.L.b200vf_element_transform_frame(self.h, C.byref(fin), C.byref(fout))

    def transform_frame_ip(self, frame):
        return self.L.b200vf_element_transform_frame_ip(self.h, C.byref(frame))

    # queued operation: BaseTransform's submit_input_buffer / generate_output pair
    def set_frames_in_flight(self, frames):
        domain = self.L.b200vf_element_set_frames_in_flight(self.h, frames)
        if domain != 0:
            raise ElementError(domain, self.message())

    def submit_input_frame(self, fin, fout=None):
        """Queue a frame (fout None = in place); FlowReturn as transform_frame."""
        return self.L.b200vf_element_submit_input_frame(self.h, C.byref(fin),
                                                        C.byref(fout) if fout is not None else None)

    def generate_output(self):
        """The oldest queued output frame, complete, once more than frames_in_flight are held;
        None = no output yet."""
        done = Frame()
        rc = self.L.b200vf_element_generate_output(self.h, C.byref(done))
        if rc < 0:
            raise ElementError(0, self.message())
        return done if rc == 1 else None

    def drain(self):
        """EOS / flush: every held frame, completed, oldest first."""
        out = []
        while True:
            done = Frame()
            rc = self.L.b200vf_element_drain(self.h, C.byref(done))
            if rc < 0:
                raise ElementError(0, self.message())
            if rc == 0:
                return out
            out.append(done)

    def transform_caps(self, direction, formats, filter_formats=None):
        """direction: 'src' or 'sink' (the pad the caps are ON); returns the other pad's formats."""
        f = None if formats is None else ",".join(formats).encode()
        flt = None if filter_formats is None else ",".join(filter_formats).encode()
        out = self.L.b200vf_element_transform_caps(self.h, 1 if direction == "src" else 0, f, flt)
        return [x for x in out.decode().split(",") if x]
