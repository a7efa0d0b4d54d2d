// Link the prebuilt in-tree library: RCB200_LIB_DIR=<repo>/range_coder_rust_b200
fn main() {
    let dir = std::env::var("RCB200_LIB_DIR").expect("set RCB200_LIB_DIR to the directory holding librcb200.so");
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=rcb200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    println!("cargo:rerun-if-env-changed=RCB200_LIB_DIR");
}
