// What the two HSV elements have in common on this side of the FFI: the library context that is
// created with the element's first frame (the reference elements have no start/stop of their own,
// so there is no earlier hook that knows a streaming thread exists) and the pad-template helper.
use b200vf_sys as ffi;
use gst_video::VideoFormat;

/// Lazily created `b200vf_ctx` of one element instance.
#[derive(Default)]
pub struct Gpu {
    ctx: Option<ffi::Context>,
}

impl Gpu {
    /// The context for `device`, created on first use and re-created when the property changed.
    pub fn get(&mut self, device: i32) -> Result<&ffi::Context, String> {
        if self.ctx.as_ref().map(|c| c.device()) != Some(device) {
            self.ctx = Some(ffi::Context::new(device)?);
        }
        Ok(self.ctx.as_ref().unwrap())
    }

    pub fn release(&mut self) {
        self.ctx = None;
    }
}

pub fn pad_template(name: &str, dir: gst::PadDirection, formats: &[VideoFormat]) -> gst::PadTemplate {
    let caps = gst_video::VideoCapsBuilder::new()
        .format_list(formats.iter().copied())
        .build();
    gst::PadTemplate::new(name, dir, gst::PadPresence::Always, &caps).unwrap()
}

/// `device` — the one property these shims add to the reference's (CUDA device index, default 0).
pub fn device_pspec() -> gst::glib::ParamSpec {
    gst::glib::ParamSpecInt::builder("device")
        .nick("Device")
        .blurb("Index of the CUDA device the frames are processed on")
        .minimum(0)
        .default_value(0)
        .mutable_ready()
        .build()
}
