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
