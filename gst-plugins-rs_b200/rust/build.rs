// build.rs fragment for video/colorlut and video/hsv: link the B200 library.
// B200VF_LIB_DIR points at the directory holding libb200vf.so.
fn main() {
    gst_plugin_version_helper::info();
    if let Ok(dir) = std::env::var("B200VF_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
    }
    println!("cargo:rustc-link-lib=dylib=b200vf");
}
