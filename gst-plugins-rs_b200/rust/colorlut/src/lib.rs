//! Plugin `colorlut` (library gstcolorlut, licence MPL-2.0) with the B200 path behind it.
//!
//! The plugin surface is that of video/colorlut/src/lib.rs of the reference: one element,
//! `colorlut`.  What is gone is `mod parser` — the .cube file is parsed inside libb200vf.so (same
//! grammar, same limits, same error texts: csrc/vf_cube_parser.cpp) when the element starts — and
//! the per-pixel loops of colorlut/imp.rs, which are the CUDA kernels behind
//! `b200vf_colorlut_process`.

use gst::glib;

mod colorlut;

/// Called by GStreamer when the shared object is loaded: registers the one element.
fn plugin_init(plugin: &gst::Plugin) -> Result<(), glib::BoolError> {
    colorlut::register(plugin)
}

gst::plugin_define!(
    colorlut,
    env!("CARGO_PKG_DESCRIPTION"),
    plugin_init,
    concat!(env!("CARGO_PKG_VERSION"), "-", env!("COMMIT_ID")),
    "MPL-2.0",
    env!("CARGO_PKG_NAME"),
    env!("CARGO_PKG_NAME"),
    env!("CARGO_PKG_REPOSITORY"),
    env!("BUILD_REL_DATE")
);
