// build.rs of smartcore with the `cuda` feature: link libsmartcore_kmeans_cuda.so (and through it the CUDA runtime).
// SMARTCORE_CUDA_LIB_DIR points at the directory that holds the library (smartcore_b200/lib of this repository, or
// wherever it was installed); without it the system search path is used.
fn main() {
    if std::env::var_os("CARGO_FEATURE_CUDA").is_none() {
        return;
    }
    if let Some(dir) = std::env::var_os("SMARTCORE_CUDA_LIB_DIR") {
        let dir = dir.to_string_lossy().into_owned();
        println!("cargo:rustc-link-search=native={}", dir);
        println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    }
    println!("cargo:rustc-link-lib=dylib=smartcore_kmeans_cuda");
    println!("cargo:rerun-if-env-changed=SMARTCORE_CUDA_LIB_DIR");
}
