// ffi.rs — Rust view of include/b200vf.h (the drop-in boundary).  Keep in sync with the
// header; `bindgen include/b200vf.h` produces the same items.
#![allow(non_camel_case_types, dead_code)]

use std::ffi::{c_char, c_int, c_void, CStr};

pub const B200VF_OK: c_int = 0;
pub const B200VF_ERR_NO_LUT: c_int = -4;
pub const B200VF_ERR_PARSE: c_int = -5;
pub const B200VF_ERR_IO: c_int = -6;
pub const B200VF_ERR_SETTINGS: c_int = -9;

pub const B200VF_MEM_HOST: u32 = 0;
pub const B200VF_MEM_DEVICE: u32 = 1;

#[repr(u32)]
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum Format {
    Rgba = 0, Rgbx = 1, Xrgb = 2, Argb = 3, Bgrx = 4, Bgra = 5, Xbgr = 6, Abgr = 7,
    Rgb = 8, Bgr = 9, Rgba64Le = 10, Rgba64Be = 11,
}

impl Format {
    pub fn from_gst(f: gst_video::VideoFormat) -> Option<Self> {
        use gst_video::VideoFormat as V;
        Some(match f {
            V::Rgba => Self::Rgba, V::Rgbx => Self::Rgbx, V::Xrgb => Self::Xrgb,
            V::Argb => Self::Argb, V::Bgrx => Self::Bgrx, V::Bgra => Self::Bgra,
            V::Xbgr => Self::Xbgr, V::Abgr => Self::Abgr, V::Rgb => Self::Rgb,
            V::Bgr => Self::Bgr, V::Rgba64Le => Self::Rgba64Le, V::Rgba64Be => Self::Rgba64Be,
            _ => return None,
        })
    }
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct b200vf_frame {
    pub data: *mut c_void,
    pub stride: i64,
    pub width: u32,
    pub height: u32,
    pub format: u32,
    pub memory: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct b200vf_hsvfilter_params {
    pub hue_shift: f32,
    pub saturation_mul: f32,
    pub saturation_off: f32,
    pub value_mul: f32,
    pub value_off: f32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct b200vf_hsvdetector_params {
    pub hue_ref: f32,
    pub hue_var: f32,
    pub saturation_ref: f32,
    pub saturation_var: f32,
    pub value_ref: f32,
    pub value_var: f32,
}

#[repr(C)]
pub struct b200vf_ctx {
    _private: [u8; 0],
}

extern "C" {
    pub fn b200vf_ctx_create(device: c_int, out: *mut *mut b200vf_ctx) -> c_int;
    pub fn b200vf_ctx_destroy(ctx: *mut b200vf_ctx);
    pub fn b200vf_last_error(ctx: *const b200vf_ctx) -> *const c_char;
    pub fn b200vf_ctx_synchronize(ctx: *mut b200vf_ctx) -> c_int;
    pub fn b200vf_colorlut_set_lut_file(ctx: *mut b200vf_ctx, location: *const c_char) -> c_int;
    pub fn b200vf_colorlut_clear_lut(ctx: *mut b200vf_ctx) -> c_int;
    pub fn b200vf_colorlut_process(
        ctx: *mut b200vf_ctx,
        in_: *const b200vf_frame,
        out: *const b200vf_frame,
    ) -> c_int;
    pub fn b200vf_hsvfilter_process(
        ctx: *mut b200vf_ctx,
        frame: *const b200vf_frame,
        params: *const b200vf_hsvfilter_params,
    ) -> c_int;
    pub fn b200vf_hsvdetector_process(
        ctx: *mut b200vf_ctx,
        in_: *const b200vf_frame,
        out: *const b200vf_frame,
        params: *const b200vf_hsvdetector_params,
    ) -> c_int;

    // Frame pools: what a `gst::BufferPool` subclass wraps for `propose_allocation` /
    // `decide_allocation` (host_pinned = 1: page-locked system memory; 0: memory:CUDAMemory).
    pub fn b200vf_pool_create(
        device: c_int,
        config: *const b200vf_pool_config,
        out: *mut *mut b200vf_pool,
    ) -> c_int;
    pub fn b200vf_pool_destroy(pool: *mut b200vf_pool);
    pub fn b200vf_pool_acquire(pool: *mut b200vf_pool, flags: u32, out: *mut b200vf_frame) -> c_int;
    pub fn b200vf_pool_release(
        pool: *mut b200vf_pool,
        frame: *const b200vf_frame,
        last_use_stream: *mut c_void,
    ) -> c_int;
    pub fn b200vf_pointer_info(p: *const c_void, memory: *mut u32, device: *mut c_int) -> c_int;
}

#[repr(C)]
pub struct b200vf_pool {
    _private: [u8; 0],
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct b200vf_pool_config {
    pub width: u32,
    pub height: u32,
    pub format: u32,
    pub min_buffers: u32,
    pub max_buffers: u32,
    pub host_pinned: u32,
}

/// Owning handle; one per element instance (created in `start`, dropped in `stop`).
pub struct Context(*mut b200vf_ctx);

// The C library is re-entrant across contexts and a context is only ever used from the
// element's streaming thread while the state mutex is held.
unsafe impl Send for Context {}

impl Context {
    pub fn new(device: i32) -> Result<Self, String> {
        let mut raw = std::ptr::null_mut();
        let rc = unsafe { b200vf_ctx_create(device, &mut raw) };
        if rc != B200VF_OK {
            return Err(last_error(std::ptr::null()));
        }
        Ok(Self(raw))
    }

    pub fn as_ptr(&self) -> *mut b200vf_ctx {
        self.0
    }

    pub fn last_error(&self) -> String {
        last_error(self.0)
    }
}

impl Drop for Context {
    fn drop(&mut self) {
        unsafe { b200vf_ctx_destroy(self.0) }
    }
}

fn last_error(ctx: *const b200vf_ctx) -> String {
    unsafe { CStr::from_ptr(b200vf_last_error(ctx)) }
        .to_string_lossy()
        .into_owned()
}

/// Plane 0 of a mapped frame as the C ABI wants it (colorlut/imp.rs:242-249).
pub fn frame_of<T>(f: &gst_video::VideoFrameRef<T>, data: *mut c_void) -> Option<b200vf_frame> {
    Some(b200vf_frame {
        data,
        stride: f.plane_stride()[0] as i64,
        width: f.width(),
        height: f.height(),
        format: Format::from_gst(f.format())? as u32,
        memory: B200VF_MEM_HOST,
    })
}
