// `hsvdetector` — GstHsvDetector: RGB-ish in, the same colours plus an alpha mask out.  Shell as in
// the reference (video/hsv/src/hsvdetector/imp.rs:20-98, 162-421): six float properties mutable in
// PLAYING, sink {RGBx,xRGB,BGRx,xBGR,RGB,BGR}, src {RGBA,ARGB,BGRA,ABGR}, NeverInPlace,
// transform_caps swapping the format lists.  transform_frame calls b200vf_hsvdetector_process; the
// sixteen closure pairs of :428-704 are two pixel layouts handed to the kernel.
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
        "hsvdetector",
        gst::DebugColorFlags::empty(),
        Some("Rust HSV-based detection filter"),
    )
});

const INPUT_FORMATS: [VideoFormat; 6] = [
    VideoFormat::Rgbx,
    VideoFormat::Xrgb,
    VideoFormat::Bgrx,
    VideoFormat::Xbgr,
    VideoFormat::Rgb,
    VideoFormat::Bgr,
];
const OUTPUT_FORMATS: [VideoFormat; 4] =
    [VideoFormat::Rgba, VideoFormat::Argb, VideoFormat::Bgra, VideoFormat::Abgr];

/// name, nick, blurb, default, (min, max) — None = the whole f32 range
const FLOAT_PROPS: [(&str, &str, &str, f32, Option<(f32, f32)>); 6] = [
    ("hue-ref", "Hue reference", "Hue reference in degrees", 0.0, None),
    (
        "hue-var",
        "Hue variation",
        "Allowed hue variation from the reference hue angle, in degrees",
        10.0,
        Some((0.0, 180.0)),
    ),
    ("saturation-ref", "Saturation reference", "Reference saturation value", 0.0, Some((0.0, 1.0))),
    (
        "saturation-var",
        "Saturation variation",
        "Allowed saturation variation from the reference value",
        0.15,
        Some((0.0, 1.0)),
    ),
    ("value-ref", "Value reference", "Reference value value", 0.0, Some((0.0, 1.0))),
    (
        "value-var",
        "Value variation",
        "Allowed value variation from the reference value",
        0.3,
        Some((0.0, 1.0)),
    ),
];

struct Settings {
    params: ffi::b200vf_hsvdetector_params,
    device: i32,
}

impl Default for Settings {
    fn default() -> Self {
        Settings {
            params: ffi::b200vf_hsvdetector_params {
                hue_ref: FLOAT_PROPS[0].3,
                hue_var: FLOAT_PROPS[1].3,
                saturation_ref: FLOAT_PROPS[2].3,
                saturation_var: FLOAT_PROPS[3].3,
                value_ref: FLOAT_PROPS[4].3,
                value_var: FLOAT_PROPS[5].3,
            },
            device: 0,
        }
    }
}

impl Settings {
    fn field(&mut self, name: &str) -> Option<&mut f32> {
        Some(match name {
            "hue-ref" => &mut self.params.hue_ref,
            "hue-var" => &mut self.params.hue_var,
            "saturation-ref" => &mut self.params.saturation_ref,
            "saturation-var" => &mut self.params.saturation_var,
            "value-ref" => &mut self.params.value_ref,
            "value-var" => &mut self.params.value_var,
            _ => return None,
        })
    }
}

#[derive(Default)]
pub struct HsvDetector {
    settings: Mutex<Settings>,
    gpu: Mutex<Gpu>,
}

#[glib::object_subclass]
impl ObjectSubclass for HsvDetector {
    const NAME: &'static str = "GstHsvDetector";
    type Type = super::HsvDetector;
    type ParentType = gst_video::VideoFilter;
}

impl ObjectImpl for HsvDetector {
    fn properties() -> &'static [glib::ParamSpec] {
        static PROPERTIES: LazyLock<Vec<glib::ParamSpec>> = LazyLock::new(|| {
            let mut props: Vec<glib::ParamSpec> = FLOAT_PROPS
                .iter()
                .map(|(name, nick, blurb, default, range)| {
                    let mut b = glib::ParamSpecFloat::builder(name)
                        .nick(nick)
                        .blurb(blurb)
                        .default_value(*default)
                        .mutable_playing();
                    if let Some((min, max)) = range {
                        b = b.minimum(*min).maximum(*max);
                    }
                    b.build()
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

impl GstObjectImpl for HsvDetector {}

impl ElementImpl for HsvDetector {
    fn metadata() -> Option<&'static gst::subclass::ElementMetadata> {
        static METADATA: LazyLock<gst::subclass::ElementMetadata> = LazyLock::new(|| {
            gst::subclass::ElementMetadata::new(
                "HSV detector",
                "Filter/Effect/Converter/Video",
                "Works within the HSV colorspace to mark positive pixels",
                "Julien Bardagi <julien.bardagi@gmail.com>",
            )
        });
        Some(&*METADATA)
    }

    fn pad_templates() -> &'static [gst::PadTemplate] {
        static TEMPLATES: LazyLock<Vec<gst::PadTemplate>> = LazyLock::new(|| {
            vec![
                pad_template("src", gst::PadDirection::Src, &OUTPUT_FORMATS),
                pad_template("sink", gst::PadDirection::Sink, &INPUT_FORMATS),
            ]
        });
        TEMPLATES.as_ref()
    }
}

impl BaseTransformImpl for HsvDetector {
    const MODE: gst_base::subclass::BaseTransformMode =
        gst_base::subclass::BaseTransformMode::NeverInPlace;
    const PASSTHROUGH_ON_SAME_CAPS: bool = false;
    const TRANSFORM_IP_ON_PASSTHROUGH: bool = false;

    /// Everything but the format carries over; the format list is the other pad's (imp.rs:386-420).
    fn transform_caps(
        &self,
        direction: gst::PadDirection,
        caps: &gst::Caps,
        filter: Option<&gst::Caps>,
    ) -> Option<gst::Caps> {
        let formats: &[VideoFormat] = match direction {
            gst::PadDirection::Src => &INPUT_FORMATS,
            _ => &OUTPUT_FORMATS,
        };
        let mut other = caps.clone();
        for s in other.make_mut().iter_mut() {
            s.set("format", gst::List::new(formats.iter().copied()));
        }
        gst::debug!(CAT, imp = self, "Transformed caps from {caps} to {other} in direction {direction:?}");
        Some(match filter {
            Some(filter) => filter.intersect_with_mode(&other, gst::CapsIntersectMode::First),
            None => other,
        })
    }

    fn stop(&self) -> Result<(), gst::ErrorMessage> {
        self.gpu.lock().unwrap().release();
        Ok(())
    }
}

impl VideoFilterImpl for HsvDetector {
    fn transform_frame(
        &self,
        in_frame: &gst_video::VideoFrameRef<&gst::BufferRef>,
        out_frame: &mut gst_video::VideoFrameRef<&mut gst::BufferRef>,
    ) -> Result<gst::FlowSuccess, gst::FlowError> {
        let (params, device) = {
            let settings = self.settings.lock().unwrap();
            (settings.params, settings.device)
        };

        let src = in_frame.plane_data(0).map_err(|_| gst::FlowError::Error)?.as_ptr();
        let fin = ffi::host_frame(in_frame, src as *mut _).ok_or(gst::FlowError::NotNegotiated)?;
        let dst = out_frame.plane_data_mut(0).map_err(|_| gst::FlowError::Error)?.as_mut_ptr();
        let fout = ffi::host_frame(out_frame, dst as *mut _).ok_or(gst::FlowError::NotNegotiated)?;

        let mut gpu = self.gpu.lock().unwrap();
        let ctx = gpu.get(device).map_err(|err| {
            gst::error!(CAT, imp = self, "CUDA device {device}: {err}");
            gst::FlowError::Error
        })?;
        ctx.hsvdetector(&fin, &fout, &params).map_err(|err| {
            gst::error!(CAT, imp = self, "hsvdetector: {err}");
            gst::FlowError::Error
        })?;
        Ok(gst::FlowSuccess::Ok)
    }
}
