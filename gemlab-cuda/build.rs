// Links libgemlab_cuda.so. Set GEMLAB_CUDA_LIB_DIR to the directory that holds it
// (in this repository: gemlab_b200/, built by `make -C gemlab_b200/csrc`).
fn main() {
    let dir = std::env::var("GEMLAB_CUDA_LIB_DIR").unwrap_or_else(|_| "../gemlab_b200".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=gemlab_cuda");
    println!("cargo:rerun-if-env-changed=GEMLAB_CUDA_LIB_DIR");
}
