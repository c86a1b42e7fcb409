// Links the prebuilt shared library (make -C ann-search-rs_b200) -- no CUDA compilation happens from cargo.
fn main() {
    let dir = std::env::var("ANNB200_LIB_DIR").unwrap_or_else(|_| "../../lib".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=annb200");
    println!("cargo:rerun-if-env-changed=ANNB200_LIB_DIR");
}
