use gst::glib;
use gst::prelude::*;

mod imp;

glib::wrapper! {
    pub struct HsvDetector(ObjectSubclass<imp::HsvDetector>)
        @extends gst_video::VideoFilter, gst_base::BaseTransform, gst::Element, gst::Object;
}

pub fn register(plugin: &gst::Plugin) -> Result<(), glib::BoolError> {
    gst::Element::register(Some(plugin), "hsvdetector", gst::Rank::NONE, HsvDetector::static_type())
}
