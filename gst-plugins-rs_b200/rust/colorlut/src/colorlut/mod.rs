//! The `colorlut` element type and its registration.
//!
//! Only the GObject shell lives on the Rust side: `imp::ColorLut` keeps the `location` property and
//! the pad templates of the reference element and hands every frame to `b200vf_colorlut_process`
//! (include/b200vf.h), which runs the LUT on the GPU.  Type hierarchy, factory name and rank have to
//! be the reference's for `gst-inspect-1.0 colorlut` and existing pipelines to see the same element.
use gst::glib;
use gst::prelude::*;

mod imp;

glib::wrapper! {
    pub struct ColorLut(ObjectSubclass<imp::ColorLut>)
        @extends gst_video::VideoFilter, gst_base::BaseTransform, gst::Element, gst::Object;
}

/// Element name, rank and type name are the reference's (docs/plugins/gst_plugins_cache.json).
pub fn register(plugin: &gst::Plugin) -> Result<(), glib::BoolError> {
    gst::Element::register(Some(plugin), "colorlut", gst::Rank::NONE, ColorLut::static_type())
}
