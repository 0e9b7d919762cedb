// Links libb200vf.so.  B200VF_LIB_DIR = the directory that holds it (this repository:
// gst-plugins-rs_b200/); without it the default linker search path is used.
fn main() {
    println!("cargo:rerun-if-env-changed=B200VF_LIB_DIR");
    if let Ok(dir) = std::env::var("B200VF_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    }
    println!("cargo:rustc-link-lib=dylib=b200vf");
}
