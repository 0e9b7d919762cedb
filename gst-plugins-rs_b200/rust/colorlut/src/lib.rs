// Plugin `colorlut` (library gstcolorlut, licence MPL-2.0) with the B200 path behind it.
//
// Same plugin surface as video/colorlut/src/lib.rs of the reference: one element, `colorlut`.
// What is gone is `mod parser` — the .cube file is parsed by libb200vf.so (same grammar, same
// error texts, csrc/vf_cube_parser.cpp) — and the per-pixel loops of colorlut/imp.rs.
use gst::glib;

mod colorlut;

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
