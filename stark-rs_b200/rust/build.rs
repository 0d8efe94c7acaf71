// Links the B200 back end.  NOTE: never compiled in this repository's CI image (no rustc); see INTEGRATION.md.
fn main() {
    let dir = std::env::var("STARK_B200_LIB_DIR").unwrap_or_else(|_| "..".into());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=stark_b200");
    println!("cargo:rerun-if-changed=../../include/stark_b200.h");
}
