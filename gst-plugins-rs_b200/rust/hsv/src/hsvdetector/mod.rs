//! The `hsvdetector` element type and its registration.
//!
//! `imp::HsvDetector` keeps the six float properties and the RGB-family -> alpha-carrying caps
//! transformation of the reference element and hands every buffer pair to
//! `b200vf_hsvdetector_process` (include/b200vf.h).  Type hierarchy, factory name and rank are the
//! reference's.
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
