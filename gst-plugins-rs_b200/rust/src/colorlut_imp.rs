// colorlut_imp.rs — replaces three bodies in video/colorlut/src/colorlut/imp.rs.
// Unchanged: Settings, properties (:69-102), metadata (:106-117), pad templates (:120-159),
// MODE = NeverInPlace (:163-166).  `parser.rs` is no longer needed by the element: the
// library parses the file with the same grammar and error texts.

use crate::ffi;
use std::ffi::CString;

#[derive(Default)]
struct State {
    ctx: Option<ffi::Context>, // was: lut: Option<CubeLut>   (imp.rs:50-53)
}

impl BaseTransformImpl for ColorLut {
    const MODE: gst_base::subclass::BaseTransformMode =
        gst_base::subclass::BaseTransformMode::NeverInPlace;
    const PASSTHROUGH_ON_SAME_CAPS: bool = false;
    const TRANSFORM_IP_ON_PASSTHROUGH: bool = false;

    // imp.rs:168-194
    fn start(&self) -> Result<(), gst::ErrorMessage> {
        let location = self.settings.lock().unwrap().location.clone().ok_or_else(|| {
            gst::error_msg!(
                gst::ResourceError::Settings,
                ["LUT file location is not configured"]
            )
        })?;

        let ctx = ffi::Context::new(0)
            .map_err(|err| gst::error_msg!(gst::ResourceError::Failed, ["CUDA context: {err}"]))?;
        let c_location = CString::new(location.clone()).unwrap();
        let rc = unsafe { ffi::b200vf_colorlut_set_lut_file(ctx.as_ptr(), c_location.as_ptr()) };
        if rc != ffi::B200VF_OK {
            // last_error() is already "Failed to parse LUT file {location}: {err}"
            return Err(gst::error_msg!(gst::ResourceError::Read, ["{}", ctx.last_error()]));
        }

        *self.state.lock().unwrap() = State { ctx: Some(ctx) };
        Ok(())
    }

    // imp.rs:196-199
    fn stop(&self) -> Result<(), gst::ErrorMessage> {
        *self.state.lock().unwrap() = State::default();
        Ok(())
    }
}

impl VideoFilterImpl for ColorLut {
    // imp.rs:203-223; the loops at :226-397 and helpers :399-543 are deleted.
    fn transform_frame(
        &self,
        in_frame: &gst_video::VideoFrameRef<&gst::BufferRef>,
        out_frame: &mut gst_video::VideoFrameRef<&mut gst::BufferRef>,
    ) -> Result<gst::FlowSuccess, gst::FlowError> {
        let state = self.state.lock().unwrap();
        let ctx = state.ctx.as_ref().ok_or_else(|| {
            gst::error!(CAT, imp = self, "No LUT configured");
            gst::FlowError::Error
        })?;

        let src = in_frame.plane_data(0).unwrap().as_ptr() as *mut _;
        let fin = ffi::frame_of(in_frame, src).ok_or(gst::FlowError::NotNegotiated)?;
        let dst = out_frame.plane_data_mut(0).unwrap().as_mut_ptr() as *mut _;
        let fout = ffi::frame_of(out_frame, dst).ok_or(gst::FlowError::NotNegotiated)?;

        // Host frames: complete (H2D → kernel → D2H) when the call returns.
        let rc = unsafe { ffi::b200vf_colorlut_process(ctx.as_ptr(), &fin, &fout) };
        if rc != ffi::B200VF_OK {
            gst::error!(CAT, imp = self, "colorlut: {}", ctx.last_error());
            return Err(gst::FlowError::Error);
        }
        Ok(gst::FlowSuccess::Ok)
    }
}
