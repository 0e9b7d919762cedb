// Plugin `hsv` (library gsthsv, licence MIT/X11) with the B200 path behind it: the elements
// `hsvfilter` and `hsvdetector` as in video/hsv/src/lib.rs of the reference.  `mod hsvutils` (the
// scalar RGB<->HSV helpers) is gone from the elements' path: the conversions run inside the kernels
// of libb200vf.so, bit-compatible with hsvutils.rs:44-198.
#![allow(clippy::non_send_fields_in_send_ty)]

use gst::glib;

mod hsvdetector;
mod hsvfilter;
mod shared;

fn plugin_init(plugin: &gst::Plugin) -> Result<(), glib::BoolError> {
    hsvfilter::register(plugin)?;
    hsvdetector::register(plugin)?;
    Ok(())
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
