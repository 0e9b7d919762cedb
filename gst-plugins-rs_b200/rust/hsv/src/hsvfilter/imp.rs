// `hsvfilter` — GstHsvFilter, an in-place GstVideoFilter.  Shell as in the reference
// (video/hsv/src/hsvfilter/imp.rs:20-75, 122-321): five float properties mutable in PLAYING, ten
// packed 8-bit formats on both pads, AlwaysInPlace.  transform_frame_ip hands the mapped frame and
// a snapshot of the settings to b200vf_hsvfilter_process; the per-pixel loop (:76-120) and the
// format ladder (:327-371) live in the library (the pixel layout is a kernel parameter there).
use gst::glib;
use gst::prelude::*;
use gst::subclass::prelude::*;
use gst_base::subclass::prelude::*;
use gst_video::subclass::prelude::*;
use gst_video::VideoFormat;

use crate::shared::{device_pspec, pad_template, Gpu};
use b200vf_sys as ffi;
use std::sync::{LazyLock, Mutex};

static CAT: LazyLock<gst::DebugCategory> = LazyLock::new(|| {
    gst::DebugCategory::new(
        "hsvfilter",
        gst::DebugColorFlags::empty(),
        Some("Rust HSV-based filter"),
    )
});

const FORMATS: [VideoFormat; 10] = [
    VideoFormat::Rgbx,
    VideoFormat::Xrgb,
    VideoFormat::Bgrx,
    VideoFormat::Xbgr,
    VideoFormat::Rgba,
    VideoFormat::Argb,
    VideoFormat::Bgra,
    VideoFormat::Abgr,
    VideoFormat::Rgb,
    VideoFormat::Bgr,
];

/// name, nick, blurb, default — all floats over the full f32 range, mutable in PLAYING
const FLOAT_PROPS: [(&str, &str, &str, f32); 5] = [
    ("hue-shift", "Hue shift", "Hue shifting in degrees", 0.0),
    (
        "saturation-mul",
        "Saturation multiplier",
        "Saturation multiplier to apply to the saturation value (before offset)",
        1.0,
    ),
    (
        "saturation-off",
        "Saturation offset",
        "Saturation offset to add to the saturation value (after multiplier)",
        0.0,
    ),
    (
        "value-mul",
        "Value multiplier",
        "Value multiplier to apply to the value (before offset)",
        1.0,
    ),
    (
        "value-off",
        "Value offset",
        "Value offset to add to the value (after multiplier)",
        0.0,
    ),
];

struct Settings {
    params: ffi::b200vf_hsvfilter_params,
    device: i32,
}

impl Default for Settings {
    fn default() -> Self {
        Settings {
            params: ffi::b200vf_hsvfilter_params {
                hue_shift: FLOAT_PROPS[0].3,
                saturation_mul: FLOAT_PROPS[1].3,
                saturation_off: FLOAT_PROPS[2].3,
                value_mul: FLOAT_PROPS[3].3,
                value_off: FLOAT_PROPS[4].3,
            },
            device: 0,
        }
    }
}

impl Settings {
    fn field(&mut self, name: &str) -> Option<&mut f32> {
        Some(match name {
            "hue-shift" => &mut self.params.hue_shift,
            "saturation-mul" => &mut self.params.saturation_mul,
            "saturation-off" => &mut self.params.saturation_off,
            "value-mul" => &mut self.params.value_mul,
            "value-off" => &mut self.params.value_off,
            _ => return None,
        })
    }
}

#[derive(Default)]
pub struct HsvFilter {
    // properties are written from application threads while frames flow (imp.rs:229-230)
    settings: Mutex<Settings>,
    gpu: Mutex<Gpu>,
}

#[glib::object_subclass]
impl ObjectSubclass for HsvFilter {
    const NAME: &'static str = "GstHsvFilter";
    type Type = super::HsvFilter;
    type ParentType = gst_video::VideoFilter;
}

impl ObjectImpl for HsvFilter {
    fn properties() -> &'static [glib::ParamSpec] {
        static PROPERTIES: LazyLock<Vec<glib::ParamSpec>> = LazyLock::new(|| {
            let mut props: Vec<glib::ParamSpec> = FLOAT_PROPS
                .iter()
                .map(|(name, nick, blurb, default)| {
                    glib::ParamSpecFloat::builder(name)
                        .nick(nick)
                        .blurb(blurb)
                        .default_value(*default)
                        .mutable_playing()
                        .build()
                })
                .collect();
            props.push(device_pspec());
            props
        });
        PROPERTIES.as_ref()
    }

    fn set_property(&self, _id: usize, value: &glib::Value, pspec: &glib::ParamSpec) {
        let mut settings = self.settings.lock().unwrap();
        if pspec.name() == "device" {
            settings.device = value.get().expect("type checked upstream");
            return;
        }
        let new: f32 = value.get().expect("type checked upstream");
        let field = settings.field(pspec.name()).unwrap_or_else(|| unimplemented!());
        gst::info!(CAT, imp = self, "Changing {} from {} to {}", pspec.name(), *field, new);
        *field = new;
    }

    fn property(&self, _id: usize, pspec: &glib::ParamSpec) -> glib::Value {
        let mut settings = self.settings.lock().unwrap();
        if pspec.name() == "device" {
            return settings.device.to_value();
        }
        settings.field(pspec.name()).unwrap_or_else(|| unimplemented!()).to_value()
    }
}

impl GstObjectImpl for HsvFilter {}

impl ElementImpl for HsvFilter {
    fn metadata() -> Option<&'static gst::subclass::ElementMetadata> {
        static METADATA: LazyLock<gst::subclass::ElementMetadata> = LazyLock::new(|| {
            gst::subclass::ElementMetadata::new(
                "HSV filter",
                "Filter/Effect/Converter/Video",
                "Works within the HSV colorspace to apply transformations to incoming frames",
                "Julien Bardagi <julien.bardagi@gmail.com>",
            )
        });
        Some(&*METADATA)
    }

    fn pad_templates() -> &'static [gst::PadTemplate] {
        static TEMPLATES: LazyLock<Vec<gst::PadTemplate>> = LazyLock::new(|| {
            vec![
                pad_template("src", gst::PadDirection::Src, &FORMATS),
                pad_template("sink", gst::PadDirection::Sink, &FORMATS),
            ]
        });
        TEMPLATES.as_ref()
    }
}

impl BaseTransformImpl for HsvFilter {
    const MODE: gst_base::subclass::BaseTransformMode =
        gst_base::subclass::BaseTransformMode::AlwaysInPlace;
    const PASSTHROUGH_ON_SAME_CAPS: bool = false;
    const TRANSFORM_IP_ON_PASSTHROUGH: bool = false;

    fn stop(&self) -> Result<(), gst::ErrorMessage> {
        self.gpu.lock().unwrap().release(); // the context and its function table go with the stream
        Ok(())
    }
}

impl VideoFilterImpl for HsvFilter {
    fn transform_frame_ip(
        &self,
        frame: &mut gst_video::VideoFrameRef<&mut gst::BufferRef>,
    ) -> Result<gst::FlowSuccess, gst::FlowError> {
        // one snapshot per frame, then the lock is gone (imp.rs:85)
        let (params, device) = {
            let settings = self.settings.lock().unwrap();
            (settings.params, settings.device)
        };

        let data = frame.plane_data_mut(0).map_err(|_| gst::FlowError::Error)?.as_mut_ptr();
        let desc = ffi::host_frame(frame, data as *mut _).ok_or(gst::FlowError::NotNegotiated)?;

        let mut gpu = self.gpu.lock().unwrap();
        let ctx = gpu.get(device).map_err(|err| {
            gst::error!(CAT, imp = self, "CUDA device {device}: {err}");
            gst::FlowError::Error
        })?;
        ctx.hsvfilter(&desc, &params).map_err(|err| {
            gst::error!(CAT, imp = self, "hsvfilter: {err}");
            gst::FlowError::Error
        })?;
        Ok(gst::FlowSuccess::Ok)
    }
}
