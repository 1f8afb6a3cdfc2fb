// Links libfootile_b200.so (built by `make -C footile_b200/csrc`).  FOOTILE_B200_LIB_DIR names the
// directory that holds it; default: ../../footile_b200 relative to this crate.
fn main() {
    let dir = std::env::var("FOOTILE_B200_LIB_DIR").unwrap_or_else(|_| {
        let here = std::env::var("CARGO_MANIFEST_DIR").unwrap();
        format!("{}/../../footile_b200", here)
    });
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=footile_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    println!("cargo:rerun-if-env-changed=FOOTILE_B200_LIB_DIR");
}
