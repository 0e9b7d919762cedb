// hsvfilter_imp.rs — replaces HsvFilter::hsv_filter (video/hsv/src/hsvfilter/imp.rs:74-121)
// and transform_frame_ip (:322-376).  Unchanged: Settings + defaults (:24-52), properties
// (:123-254), metadata (:259-271), pad templates (:274-312), MODE = AlwaysInPlace (:316-319).
// hsvutils.rs is no longer used by the element.

use crate::ffi;
use std::sync::Mutex;

#[derive(Default)]
pub struct HsvFilter {
    settings: Mutex<Settings>,
    ctx: Mutex<Option<ffi::Context>>, // new: created lazily on the streaming thread
}

impl BaseTransformImpl for HsvFilter {
    const MODE: gst_base::subclass::BaseTransformMode =
        gst_base::subclass::BaseTransformMode::AlwaysInPlace;
    const PASSTHROUGH_ON_SAME_CAPS: bool = false;
    const TRANSFORM_IP_ON_PASSTHROUGH: bool = false;

    fn stop(&self) -> Result<(), gst::ErrorMessage> {
        *self.ctx.lock().unwrap() = None;
        Ok(())
    }
}

impl VideoFilterImpl for HsvFilter {
    fn transform_frame_ip(
        &self,
        frame: &mut gst_video::VideoFrameRef<&mut gst::BufferRef>,
    ) -> Result<gst::FlowSuccess, gst::FlowError> {
        // imp.rs:85 — one snapshot of the settings per frame
        let s = *self.settings.lock().unwrap();
        let params = ffi::b200vf_hsvfilter_params {
            hue_shift: s.hue_shift,
            saturation_mul: s.saturation_mul,
            saturation_off: s.saturation_off,
            value_mul: s.value_mul,
            value_off: s.value_off,
        };

        let mut guard = self.ctx.lock().unwrap();
        if guard.is_none() {
            *guard = Some(ffi::Context::new(0).map_err(|err| {
                gst::error!(CAT, imp = self, "CUDA context: {err}");
                gst::FlowError::Error
            })?);
        }
        let ctx = guard.as_ref().unwrap();

        // The ten formats of imp.rs:327-371 are selected inside the library from `format`.
        let data = frame.plane_data_mut(0).unwrap().as_mut_ptr() as *mut _;
        let f = ffi::frame_of(frame, data).ok_or(gst::FlowError::NotNegotiated)?;
        let rc = unsafe { ffi::b200vf_hsvfilter_process(ctx.as_ptr(), &f, &params) };
        if rc != ffi::B200VF_OK {
            gst::error!(CAT, imp = self, "hsvfilter: {}", ctx.last_error());
            return Err(gst::FlowError::Error);
        }
        Ok(gst::FlowSuccess::Ok)
    }
}
