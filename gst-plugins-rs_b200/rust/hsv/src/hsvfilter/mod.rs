//! The `hsvfilter` element type and its registration.
//!
//! `imp::HsvFilter` keeps the five float properties of the reference element (live-settable, one
//! snapshot per buffer) and hands every buffer, in place, to `b200vf_hsvfilter_process`
//! (include/b200vf.h).  Type hierarchy, factory name and rank are the reference's.
use gst::glib;
use gst::prelude::*;

mod imp;

glib::wrapper! {
    pub struct HsvFilter(ObjectSubclass<imp::HsvFilter>)
        @extends gst_video::VideoFilter, gst_base::BaseTransform, gst::Element, gst::Object;
}

pub fn register(plugin: &gst::Plugin) -> Result<(), glib::BoolError> {
    gst::Element::register(Some(plugin), "hsvfilter", gst::Rank::NONE, HsvFilter::static_type())
}
