// Links the CUDA path-tracing library (include/mrt.h).  MRT_LIB_DIR = directory of libmrt.so.
fn main() {
    let dir = std::env::var("MRT_LIB_DIR").unwrap_or_else(|_| "/usr/local/lib".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=mrt");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    println!("cargo:rerun-if-env-changed=MRT_LIB_DIR");
}
