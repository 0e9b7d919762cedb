// b200vf-sys — Rust view of include/b200vf.h, the C ABI of libb200vf.so (the B200-native colour
// transform path).  Part 1 declares every exported symbol exactly as the header does
// (tests/test_rust_shim.py diffs the two); part 2 holds the few safe wrappers the elements use.
#![allow(non_camel_case_types, clippy::missing_safety_doc)]

use std::ffi::{c_char, c_int, c_void, CStr, CString};

// ---- status codes (b200vf_status) ----------------------------------------------------------------
pub const B200VF_OK: c_int = 0;
pub const B200VF_ERR_INVALID_ARG: c_int = -1;
pub const B200VF_ERR_UNSUPPORTED_FORMAT: c_int = -2;
pub const B200VF_ERR_CUDA: c_int = -3;
pub const B200VF_ERR_NO_LUT: c_int = -4;
pub const B200VF_ERR_PARSE: c_int = -5;
pub const B200VF_ERR_IO: c_int = -6;
pub const B200VF_ERR_NO_DEVICE: c_int = -7;
pub const B200VF_ERR_NOMEM: c_int = -8;
pub const B200VF_ERR_SETTINGS: c_int = -9;

// ---- b200vf_format / b200vf_memory ----------------------------------------------------------------
pub const B200VF_FORMAT_RGBA: u32 = 0;
pub const B200VF_FORMAT_RGBX: u32 = 1;
pub const B200VF_FORMAT_XRGB: u32 = 2;
pub const B200VF_FORMAT_ARGB: u32 = 3;
pub const B200VF_FORMAT_BGRX: u32 = 4;
pub const B200VF_FORMAT_BGRA: u32 = 5;
pub const B200VF_FORMAT_XBGR: u32 = 6;
pub const B200VF_FORMAT_ABGR: u32 = 7;
pub const B200VF_FORMAT_RGB: u32 = 8;
pub const B200VF_FORMAT_BGR: u32 = 9;
pub const B200VF_FORMAT_RGBA64_LE: u32 = 10;
pub const B200VF_FORMAT_RGBA64_BE: u32 = 11;
pub const B200VF_MEM_HOST: u32 = 0;
pub const B200VF_MEM_DEVICE: u32 = 1;
pub const B200VF_LUT_1D: u32 = 1;
pub const B200VF_LUT_3D: u32 = 3;
pub const B200VF_POOL_DONTWAIT: u32 = 1;

#[repr(C)]
pub struct b200vf_ctx {
    _opaque: [u8; 0],
}
#[repr(C)]
pub struct b200vf_group {
    _opaque: [u8; 0],
}
#[repr(C)]
pub struct b200vf_pool {
    _opaque: [u8; 0],
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct b200vf_frame {
    pub data: *mut c_void,
    pub stride: i64,
    pub width: u32,
    pub height: u32,
    pub format: u32,
    pub memory: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct b200vf_stats {
    pub kernel_launches: u64,
    pub frames: u64,
    pub h2d_bytes: u64,
    pub d2h_bytes: u64,
}

#[repr(C)]
#[derive(Debug)]
pub struct b200vf_cube {
    pub kind: u32,
    pub size: u32,
    pub domain_scale: [f32; 3],
    pub domain_offset: [f32; 3],
    pub data: *mut f32,
    pub n_floats: usize,
}

/// Property snapshot of `hsvfilter`; defaults as in the reference (hsvfilter/imp.rs:25-29).
#[repr(C)]
#[derive(Clone, Copy, Debug, PartialEq)]
pub struct b200vf_hsvfilter_params {
    pub hue_shift: f32,
    pub saturation_mul: f32,
    pub saturation_off: f32,
    pub value_mul: f32,
    pub value_off: f32,
}

/// Property snapshot of `hsvdetector`; defaults as in the reference (hsvdetector/imp.rs:26-31).
#[repr(C)]
#[derive(Clone, Copy, Debug, PartialEq)]
pub struct b200vf_hsvdetector_params {
    pub hue_ref: f32,
    pub hue_var: f32,
    pub saturation_ref: f32,
    pub saturation_var: f32,
    pub value_ref: f32,
    pub value_var: f32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct b200vf_pool_config {
    pub width: u32,
    pub height: u32,
    pub format: u32,
    pub min_buffers: u32,
    pub max_buffers: u32,
    pub host_pinned: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct b200vf_pool_stats {
    pub allocated: u32,
    pub outstanding: u32,
    pub frame_bytes: u64,
    pub stride: i64,
}

extern "C" {
    pub fn b200vf_version() -> *const c_char;
    pub fn b200vf_status_string(status: c_int) -> *const c_char;
    pub fn b200vf_device_count(count: *mut c_int) -> c_int;
    pub fn b200vf_format_bytes_per_pixel(format: u32) -> u32;
    pub fn b200vf_format_name(format: u32) -> *const c_char;
    pub fn b200vf_format_from_name(name: *const c_char) -> c_int;
    pub fn b200vf_ctx_create(device: c_int, out: *mut *mut b200vf_ctx) -> c_int;
    pub fn b200vf_ctx_destroy(ctx: *mut b200vf_ctx);
    pub fn b200vf_last_error(ctx: *const b200vf_ctx) -> *const c_char;
    pub fn b200vf_ctx_device(ctx: *const b200vf_ctx) -> c_int;
    pub fn b200vf_ctx_synchronize(ctx: *mut b200vf_ctx) -> c_int;
    pub fn b200vf_ctx_get_stream(ctx: *const b200vf_ctx) -> *mut c_void;
    pub fn b200vf_ctx_set_stream(ctx: *mut b200vf_ctx, cuda_stream: *mut c_void) -> c_int;
    pub fn b200vf_ctx_wait_for(ctx: *mut b200vf_ctx, upstream: *mut b200vf_ctx) -> c_int;
    pub fn b200vf_ctx_set_option(ctx: *mut b200vf_ctx, key: *const c_char, value: i64) -> c_int;
    pub fn b200vf_ctx_get_option(ctx: *const b200vf_ctx, key: *const c_char, value: *mut i64) -> c_int;
    pub fn b200vf_ctx_host_ticket(ctx: *const b200vf_ctx) -> u64;
    pub fn b200vf_ctx_host_wait(ctx: *mut b200vf_ctx, ticket: u64) -> c_int;
    pub fn b200vf_ctx_get_stats(ctx: *const b200vf_ctx, out: *mut b200vf_stats) -> c_int;
    pub fn b200vf_ctx_reset_stats(ctx: *mut b200vf_ctx) -> c_int;
    pub fn b200vf_host_alloc(bytes: usize, out: *mut *mut c_void) -> c_int;
    pub fn b200vf_host_free(p: *mut c_void) -> c_int;
    pub fn b200vf_host_is_pinned(p: *const c_void) -> c_int;
    pub fn b200vf_ctx_host_memory_released(ctx: *mut b200vf_ctx, p: *const c_void, bytes: usize) -> c_int;
    pub fn b200vf_device_alloc(ctx: *mut b200vf_ctx, bytes: usize, out: *mut *mut c_void) -> c_int;
    pub fn b200vf_device_free(ctx: *mut b200vf_ctx, p: *mut c_void) -> c_int;
    pub fn b200vf_memcpy(ctx: *mut b200vf_ctx, dst: *mut c_void, src: *const c_void, bytes: usize, kind: c_int) -> c_int;
    pub fn b200vf_cube_parse(text: *const c_char, len: usize, out: *mut b200vf_cube, err: *mut c_char, errlen: usize) -> c_int;
    pub fn b200vf_cube_parse_file(path: *const c_char, out: *mut b200vf_cube, err: *mut c_char, errlen: usize) -> c_int;
    pub fn b200vf_cube_free(cube: *mut b200vf_cube);
    pub fn b200vf_colorlut_set_lut(ctx: *mut b200vf_ctx, kind: u32, size: u32, data: *const f32, domain_scale: *const f32, domain_offset: *const f32) -> c_int;
    pub fn b200vf_colorlut_set_lut_file(ctx: *mut b200vf_ctx, location: *const c_char) -> c_int;
    pub fn b200vf_colorlut_clear_lut(ctx: *mut b200vf_ctx) -> c_int;
    pub fn b200vf_colorlut_process(ctx: *mut b200vf_ctx, in_: *const b200vf_frame, out: *const b200vf_frame) -> c_int;
    pub fn b200vf_colorlut_process_batch(ctx: *mut b200vf_ctx, in_: *const b200vf_frame, out: *const b200vf_frame, n_frames: usize) -> c_int;
    pub fn b200vf_colorlut_convert_process_batch(ctx: *mut b200vf_ctx, in_: *const b200vf_frame, out: *const b200vf_frame, n_frames: usize) -> c_int;
    pub fn b200vf_hsvfilter_process(ctx: *mut b200vf_ctx, frame: *const b200vf_frame, params: *const b200vf_hsvfilter_params) -> c_int;
    pub fn b200vf_hsvfilter_process_batch(ctx: *mut b200vf_ctx, frames: *const b200vf_frame, n_frames: usize, params: *const b200vf_hsvfilter_params) -> c_int;
    pub fn b200vf_hsvdetector_process(ctx: *mut b200vf_ctx, in_: *const b200vf_frame, out: *const b200vf_frame, params: *const b200vf_hsvdetector_params) -> c_int;
    pub fn b200vf_hsvdetector_process_batch(ctx: *mut b200vf_ctx, in_: *const b200vf_frame, out: *const b200vf_frame, n_frames: usize, params: *const b200vf_hsvdetector_params) -> c_int;
    pub fn b200vf_chain_lut_hsv_process_batch(ctx: *mut b200vf_ctx, in_: *const b200vf_frame, out: *const b200vf_frame, n_frames: usize, params: *const b200vf_hsvfilter_params) -> c_int;
    pub fn b200vf_group_create(devices: *const c_int, n_devices: usize, out: *mut *mut b200vf_group) -> c_int;
    pub fn b200vf_group_destroy(group: *mut b200vf_group);
    pub fn b200vf_group_size(group: *const b200vf_group) -> usize;
    pub fn b200vf_group_ctx(group: *mut b200vf_group, member: usize) -> *mut b200vf_ctx;
    pub fn b200vf_group_last_error(group: *const b200vf_group) -> *const c_char;
    pub fn b200vf_group_set_option(group: *mut b200vf_group, key: *const c_char, value: i64) -> c_int;
    pub fn b200vf_group_synchronize(group: *mut b200vf_group) -> c_int;
    pub fn b200vf_group_colorlut_set_lut(group: *mut b200vf_group, kind: u32, size: u32, data: *const f32, domain_scale: *const f32, domain_offset: *const f32) -> c_int;
    pub fn b200vf_group_colorlut_set_lut_file(group: *mut b200vf_group, location: *const c_char) -> c_int;
    pub fn b200vf_group_colorlut_clear_lut(group: *mut b200vf_group) -> c_int;
    pub fn b200vf_group_colorlut_process_batch(group: *mut b200vf_group, in_: *const b200vf_frame, out: *const b200vf_frame, n_frames: usize) -> c_int;
    pub fn b200vf_group_hsvfilter_process_batch(group: *mut b200vf_group, frames: *const b200vf_frame, n_frames: usize, params: *const b200vf_hsvfilter_params) -> c_int;
    pub fn b200vf_group_hsvdetector_process_batch(group: *mut b200vf_group, in_: *const b200vf_frame, out: *const b200vf_frame, n_frames: usize, params: *const b200vf_hsvdetector_params) -> c_int;
    pub fn b200vf_group_chain_lut_hsv_process_batch(group: *mut b200vf_group, in_: *const b200vf_frame, out: *const b200vf_frame, n_frames: usize, params: *const b200vf_hsvfilter_params) -> c_int;
    pub fn b200vf_pool_create(device: c_int, config: *const b200vf_pool_config, out: *mut *mut b200vf_pool) -> c_int;
    pub fn b200vf_pool_destroy(pool: *mut b200vf_pool);
    pub fn b200vf_pool_acquire(pool: *mut b200vf_pool, flags: u32, out: *mut b200vf_frame) -> c_int;
    pub fn b200vf_pool_release(pool: *mut b200vf_pool, frame: *const b200vf_frame, last_use_stream: *mut c_void) -> c_int;
    pub fn b200vf_pool_release_after(pool: *mut b200vf_pool, frame: *const b200vf_frame, last_user: *const b200vf_ctx) -> c_int;
    pub fn b200vf_pool_get_stats(pool: *mut b200vf_pool, out: *mut b200vf_pool_stats) -> c_int;
    pub fn b200vf_pool_device(pool: *const b200vf_pool) -> c_int;
    pub fn b200vf_pointer_info(p: *const c_void, memory: *mut u32, device: *mut c_int) -> c_int;
    pub fn b200vf_debug_table_indices(colours: *const u32, n: usize, out: *mut u32) -> c_int;
    pub fn b200vf_debug_hsv_from_rgb(ctx: *mut b200vf_ctx, rgba_device: *const c_void, n_pixels: usize, hsv_device: *mut f32) -> c_int;
}

// =====================================================================================================
// safe wrappers
// =====================================================================================================

/// GstVideoFormat -> b200vf_format for the formats of the three elements' caps.
pub fn format_from_gst(f: gst_video::VideoFormat) -> Option<u32> {
    use gst_video::VideoFormat as V;
    Some(match f {
        V::Rgba => B200VF_FORMAT_RGBA,
        V::Rgbx => B200VF_FORMAT_RGBX,
        V::Xrgb => B200VF_FORMAT_XRGB,
        V::Argb => B200VF_FORMAT_ARGB,
        V::Bgrx => B200VF_FORMAT_BGRX,
        V::Bgra => B200VF_FORMAT_BGRA,
        V::Xbgr => B200VF_FORMAT_XBGR,
        V::Abgr => B200VF_FORMAT_ABGR,
        V::Rgb => B200VF_FORMAT_RGB,
        V::Bgr => B200VF_FORMAT_BGR,
        V::Rgba64Le => B200VF_FORMAT_RGBA64_LE,
        V::Rgba64Be => B200VF_FORMAT_RGBA64_BE,
        _ => return None,
    })
}

/// Plane 0 of a mapped system-memory frame as the library sees it: pointer, stride, size, format.
/// Exactly what the reference loops read (colorlut/imp.rs:242-249, hsvfilter/imp.rs:89-97).
pub fn host_frame<T>(frame: &gst_video::VideoFrameRef<T>, data: *mut c_void) -> Option<b200vf_frame> {
    Some(b200vf_frame {
        data,
        stride: frame.plane_stride()[0] as i64,
        width: frame.width(),
        height: frame.height(),
        format: format_from_gst(frame.format())?,
        memory: B200VF_MEM_HOST,
    })
}

/// One `b200vf_ctx`: created in `start`, dropped in `stop`.  Single-caller, like the streaming
/// thread that owns it; the elements keep it behind their state mutex.
pub struct Context(*mut b200vf_ctx);

// The library allows a context to move between threads as long as calls do not overlap.
unsafe impl Send for Context {}

impl Context {
    pub fn new(device: i32) -> Result<Self, String> {
        let mut raw = std::ptr::null_mut();
        let rc = unsafe { b200vf_ctx_create(device as c_int, &mut raw) };
        if rc != B200VF_OK {
            return Err(last_error(std::ptr::null()));
        }
        Ok(Context(raw))
    }

    pub fn as_ptr(&self) -> *mut b200vf_ctx {
        self.0
    }

    pub fn device(&self) -> i32 {
        unsafe { b200vf_ctx_device(self.0) as i32 }
    }

    pub fn last_error(&self) -> String {
        last_error(self.0)
    }

    fn check(&self, rc: c_int) -> Result<(), String> {
        if rc == B200VF_OK {
            Ok(())
        } else {
            Err(self.last_error())
        }
    }

    pub fn set_option(&self, key: &str, value: i64) -> Result<(), String> {
        let key = CString::new(key).map_err(|e| e.to_string())?;
        self.check(unsafe { b200vf_ctx_set_option(self.0, key.as_ptr(), value) })
    }

    /// Ticket of the most recent call on system-memory frames (`"host.async"`, include/b200vf.h).
    pub fn host_ticket(&self) -> u64 {
        unsafe { b200vf_ctx_host_ticket(self.0) }
    }

    /// Blocks until the call with this ticket, and every earlier one, is complete.
    pub fn host_wait(&self, ticket: u64) -> Result<(), String> {
        self.check(unsafe { b200vf_ctx_host_wait(self.0, ticket) })
    }

    /// `start` of colorlut: parse + upload.  Returns the status too so the caller can tell
    /// `ResourceError::Settings` / `Read` apart like the reference does (colorlut/imp.rs:175-187).
    pub fn set_lut_file(&self, location: &str) -> Result<(), (c_int, String)> {
        let loc = CString::new(location).map_err(|e| (B200VF_ERR_INVALID_ARG, e.to_string()))?;
        let rc = unsafe { b200vf_colorlut_set_lut_file(self.0, loc.as_ptr()) };
        if rc == B200VF_OK {
            Ok(())
        } else {
            Err((rc, self.last_error()))
        }
    }

    pub fn colorlut(&self, fin: &b200vf_frame, fout: &b200vf_frame) -> Result<(), String> {
        self.check(unsafe { b200vf_colorlut_process(self.0, fin, fout) })
    }

    pub fn hsvfilter(&self, frame: &b200vf_frame, p: &b200vf_hsvfilter_params) -> Result<(), String> {
        self.check(unsafe { b200vf_hsvfilter_process(self.0, frame, p) })
    }

    pub fn hsvdetector(
        &self,
        fin: &b200vf_frame,
        fout: &b200vf_frame,
        p: &b200vf_hsvdetector_params,
    ) -> Result<(), String> {
        self.check(unsafe { b200vf_hsvdetector_process(self.0, fin, fout, p) })
    }

    /// With "host.register" = 1: upstream memory is about to be freed (call from a destroy notify
    /// on the GstMemory).
    pub fn host_memory_released(&self, p: *const c_void, bytes: usize) {
        unsafe { b200vf_ctx_host_memory_released(self.0, p, bytes) };
    }
}

impl Drop for Context {
    fn drop(&mut self) {
        unsafe { b200vf_ctx_destroy(self.0) }
    }
}

fn last_error(ctx: *const b200vf_ctx) -> String {
    unsafe {
        let p = b200vf_last_error(ctx);
        if p.is_null() {
            String::new()
        } else {
            CStr::from_ptr(p as *const c_char).to_string_lossy().into_owned()
        }
    }
}
