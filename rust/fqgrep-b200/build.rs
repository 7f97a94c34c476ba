// Links the prebuilt CUDA library (python -m fqgrep_b200.build runs nvcc for sm_100a).
fn main() {
    let dir = std::env::var("FQGREP_B200_LIB_DIR").expect("set FQGREP_B200_LIB_DIR to <repo>/fqgrep_b200");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=fqgrep_b200");
    println!("cargo:rerun-if-env-changed=FQGREP_B200_LIB_DIR");
}
