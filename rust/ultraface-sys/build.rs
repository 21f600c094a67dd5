// Links against the C-ABI library built by `python -m infercam_onnx_b200.build`.
fn main() {
    let dir = std::env::var("ULTRAFACE_B200_LIB_DIR").unwrap_or_else(|_| "../../infercam_onnx_b200".into());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=ultraface_b200");
}
