// build.rs — links libmotivate_b200.so (built by `make -C motivate_b200/csrc`, nvcc, sm_100a only).
// MOTIVATE_B200_LIB_DIR = directory that holds the shared object (default: ../../motivate_b200 of this repo).
use std::env;
use std::path::PathBuf;

fn main() {
    let dir = env::var("MOTIVATE_B200_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../motivate_b200")
    });
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=motivate_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
    println!("cargo:rerun-if-env-changed=MOTIVATE_B200_LIB_DIR");
    println!("cargo:rerun-if-changed=../../include/motivate_b200.h");
}
