// hsvdetector_imp.rs — replaces HsvDetector::hsv_detect
// (video/hsv/src/hsvdetector/imp.rs:98-161) and the 16-way closure dispatch of
// transform_frame (:422-707).  Unchanged: Settings + defaults (:25-56), format lists
// (:78-96), properties (:163-323), metadata, pad templates, MODE = NeverInPlace,
// transform_caps (:386-419).

use crate::ffi;
use std::sync::Mutex;

#[derive(Default)]
pub struct HsvDetector {
    settings: Mutex<Settings>,
    ctx: Mutex<Option<ffi::Context>>,
}

impl VideoFilterImpl for HsvDetector {
    fn transform_frame(
        &self,
        in_frame: &gst_video::VideoFrameRef<&gst::BufferRef>,
        out_frame: &mut gst_video::VideoFrameRef<&mut gst::BufferRef>,
    ) -> Result<gst::FlowSuccess, gst::FlowError> {
        let s = *self.settings.lock().unwrap(); // imp.rs:110
        let params = ffi::b200vf_hsvdetector_params {
            hue_ref: s.hue_ref,
            hue_var: s.hue_var,
            saturation_ref: s.saturation_ref,
            saturation_var: s.saturation_var,
            value_ref: s.value_ref,
            value_var: s.value_var,
        };

        let mut guard = self.ctx.lock().unwrap();
        if guard.is_none() {
            *guard = Some(ffi::Context::new(0).map_err(|err| {
                gst::error!(CAT, imp = self, "CUDA context: {err}");
                gst::FlowError::Error
            })?);
        }
        let ctx = guard.as_ref().unwrap();

        let src = in_frame.plane_data(0).unwrap().as_ptr() as *mut _;
        let fin = ffi::frame_of(in_frame, src).ok_or(gst::FlowError::NotNegotiated)?;
        let dst = out_frame.plane_data_mut(0).unwrap().as_mut_ptr() as *mut _;
        let fout = ffi::frame_of(out_frame, dst).ok_or(gst::FlowError::NotNegotiated)?;

        // (in format, out format) picks one of the 24 byte mappings inside the library.
        let rc = unsafe { ffi::b200vf_hsvdetector_process(ctx.as_ptr(), &fin, &fout, &params) };
        if rc != ffi::B200VF_OK {
            gst::error!(CAT, imp = self, "hsvdetector: {}", ctx.last_error());
            return Err(gst::FlowError::Error);
        }
        Ok(gst::FlowSuccess::Ok)
    }
}
