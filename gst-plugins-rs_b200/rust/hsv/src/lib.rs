//! Plugin `hsv` (library gsthsv, licence MIT/X11) with the B200 path behind it.
//!
//! The plugin surface is that of video/hsv/src/lib.rs of the reference — two elements, `hsvfilter`
//! and `hsvdetector`, registered in that order.  What is gone is `mod hsvutils`: the scalar
//! RGB <-> HSV helpers are no longer on the elements' path, the conversions run inside the kernels
//! of libb200vf.so, bit-compatible with hsvutils.rs:44-198.  `shared` holds what the two elements
//! have in common on this side: the lazily created device context and the pad-template helper.

use gst::glib;

mod hsvdetector;
mod hsvfilter;
mod shared;

type Registrar = fn(&gst::Plugin) -> Result<(), glib::BoolError>;

/// The elements of this plugin, in the reference's registration order.
const ELEMENTS: [Registrar; 2] = [hsvfilter::register, hsvdetector::register];

fn plugin_init(plugin: &gst::Plugin) -> Result<(), glib::BoolError> {
    ELEMENTS.iter().try_for_each(|register| register(plugin))
}

gst::plugin_define!(
    hsv,
    env!("CARGO_PKG_DESCRIPTION"),
    plugin_init,
    concat!(env!("CARGO_PKG_VERSION"), "-", env!("COMMIT_ID")),
    "MIT/X11",
    env!("CARGO_PKG_NAME"),
    env!("CARGO_PKG_NAME"),
    env!("CARGO_PKG_REPOSITORY"),
    env!("BUILD_REL_DATE")
);
