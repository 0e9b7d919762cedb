// `colorlut` — GstColorLut, a GstVideoFilter.  The GObject shell (type name, `location` property,
// metadata, pad templates, NeverInPlace mode, start/stop semantics and error domains) is the
// reference's, video/colorlut/src/colorlut/imp.rs:45-223; the work is done by libb200vf.so:
//
//   start            -> b200vf_ctx_create + b200vf_colorlut_set_lut_file   (was CubeLut::parse_file)
//   transform_frame  -> b200vf_colorlut_process                            (was transform_rgba* loops)
//   stop             -> b200vf_ctx_destroy
//
// One addition: the `device` property (CUDA device index, default 0 = what a single-GPU box has),
// mutable in READY like `location`.
use gst::glib;
use gst::prelude::*;
use gst::subclass::prelude::*;
use gst_base::subclass::prelude::*;
use gst_video::subclass::prelude::*;
use gst_video::VideoFormat;

use b200vf_sys as ffi;
use std::sync::{LazyLock, Mutex};

static CAT: LazyLock<gst::DebugCategory> = LazyLock::new(|| {
    gst::DebugCategory::new("colorlut", gst::DebugColorFlags::empty(), Some("Color LUT"))
});

struct Settings {
    location: Option<String>,
    device: i32,
}

impl Default for Settings {
    fn default() -> Self {
        Settings { location: None, device: 0 }
    }
}

/// What `start` builds and `stop` drops: the library context holding the uploaded LUT.
#[derive(Default)]
struct State {
    ctx: Option<ffi::Context>,
}

#[derive(Default)]
pub struct ColorLut {
    settings: Mutex<Settings>,
    state: Mutex<State>,
}

#[glib::object_subclass]
impl ObjectSubclass for ColorLut {
    const NAME: &'static str = "GstColorLut";
    type Type = super::ColorLut;
    type ParentType = gst_video::VideoFilter;
}

impl ObjectImpl for ColorLut {
    fn properties() -> &'static [glib::ParamSpec] {
        static PROPERTIES: LazyLock<Vec<glib::ParamSpec>> = LazyLock::new(|| {
            vec![
                glib::ParamSpecString::builder("location")
                    .nick("Location")
                    .blurb("Location of the LUT file to read from")
                    .mutable_ready()
                    .build(),
                glib::ParamSpecInt::builder("device")
                    .nick("Device")
                    .blurb("Index of the CUDA device the LUT is applied on")
                    .minimum(0)
                    .default_value(0)
                    .mutable_ready()
                    .build(),
            ]
        });
        PROPERTIES.as_ref()
    }

    fn set_property(&self, _id: usize, value: &glib::Value, pspec: &glib::ParamSpec) {
        let mut settings = self.settings.lock().unwrap();
        match pspec.name() {
            "location" => settings.location = value.get().expect("type checked upstream"),
            "device" => settings.device = value.get().expect("type checked upstream"),
            _ => unimplemented!(),
        }
    }

    fn property(&self, _id: usize, pspec: &glib::ParamSpec) -> glib::Value {
        let settings = self.settings.lock().unwrap();
        match pspec.name() {
            "location" => settings.location.to_value(),
            "device" => settings.device.to_value(),
            _ => unimplemented!(),
        }
    }
}

impl GstObjectImpl for ColorLut {}

/// `video/x-raw` templates for both pads over one format list.
fn templates(formats: &[VideoFormat]) -> Vec<gst::PadTemplate> {
    let caps = gst_video::VideoCapsBuilder::new()
        .format_list(formats.iter().copied())
        .build();
    [("sink", gst::PadDirection::Sink), ("src", gst::PadDirection::Src)]
        .into_iter()
        .map(|(name, dir)| gst::PadTemplate::new(name, dir, gst::PadPresence::Always, &caps).unwrap())
        .collect()
}

impl ElementImpl for ColorLut {
    fn metadata() -> Option<&'static gst::subclass::ElementMetadata> {
        static METADATA: LazyLock<gst::subclass::ElementMetadata> = LazyLock::new(|| {
            gst::subclass::ElementMetadata::new(
                "Color LUT",
                "Filter/Effect/Video",
                "Apply color lookup table",
                "Seungha Yang <seungha@centricular.com>",
            )
        });
        Some(&*METADATA)
    }

    fn pad_templates() -> &'static [gst::PadTemplate] {
        // native-endian 16-bit first, as in the reference (imp.rs:122-134)
        static TEMPLATES: LazyLock<Vec<gst::PadTemplate>> = LazyLock::new(|| {
            if cfg!(target_endian = "big") {
                templates(&[VideoFormat::Rgba64Be, VideoFormat::Rgba64Le, VideoFormat::Rgba])
            } else {
                templates(&[VideoFormat::Rgba64Le, VideoFormat::Rgba64Be, VideoFormat::Rgba])
            }
        });
        TEMPLATES.as_ref()
    }
}

impl BaseTransformImpl for ColorLut {
    const MODE: gst_base::subclass::BaseTransformMode =
        gst_base::subclass::BaseTransformMode::NeverInPlace;
    const PASSTHROUGH_ON_SAME_CAPS: bool = false;
    const TRANSFORM_IP_ON_PASSTHROUGH: bool = false;

    fn start(&self) -> Result<(), gst::ErrorMessage> {
        let (location, device) = {
            let settings = self.settings.lock().unwrap();
            (settings.location.clone(), settings.device)
        };
        // reference: ResourceError::Settings when `location` is unset (imp.rs:175-180)
        let location = location.ok_or_else(|| {
            gst::error_msg!(gst::ResourceError::Settings, ["LUT file location is not configured"])
        })?;

        let ctx = ffi::Context::new(device).map_err(|err| {
            gst::error_msg!(gst::ResourceError::OpenRead, ["CUDA device {device}: {err}"])
        })?;
        // reference: ResourceError::Read with "Failed to parse LUT file {location}: {err}"
        // (imp.rs:182-187); the library composes the very same text
        ctx.set_lut_file(&location)
            .map_err(|(_status, msg)| gst::error_msg!(gst::ResourceError::Read, ["{msg}"]))?;

        gst::debug!(CAT, imp = self, "LUT {location} uploaded to CUDA device {device}");
        *self.state.lock().unwrap() = State { ctx: Some(ctx) };
        Ok(())
    }

    fn stop(&self) -> Result<(), gst::ErrorMessage> {
        *self.state.lock().unwrap() = State::default(); // drops the context and every table it held
        Ok(())
    }
}

impl VideoFilterImpl for ColorLut {
    fn transform_frame(
        &self,
        in_frame: &gst_video::VideoFrameRef<&gst::BufferRef>,
        out_frame: &mut gst_video::VideoFrameRef<&mut gst::BufferRef>,
    ) -> Result<gst::FlowSuccess, gst::FlowError> {
        let state = self.state.lock().unwrap();
        let Some(ctx) = state.ctx.as_ref() else {
            gst::error!(CAT, imp = self, "No LUT configured"); // imp.rs:210-213
            return Err(gst::FlowError::Error);
        };

        let src = in_frame.plane_data(0).map_err(|_| gst::FlowError::Error)?.as_ptr();
        let fin = ffi::host_frame(in_frame, src as *mut _).ok_or(gst::FlowError::NotNegotiated)?;
        let dst = out_frame.plane_data_mut(0).map_err(|_| gst::FlowError::Error)?.as_mut_ptr();
        let fout = ffi::host_frame(out_frame, dst as *mut _).ok_or(gst::FlowError::NotNegotiated)?;

        // system-memory frames: H2D -> kernel -> D2H are complete when the call returns
        ctx.colorlut(&fin, &fout).map_err(|err| {
            gst::error!(CAT, imp = self, "colorlut: {err}");
            gst::FlowError::Error
        })?;
        Ok(gst::FlowSuccess::Ok)
    }
}
