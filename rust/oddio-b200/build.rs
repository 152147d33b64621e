// Links liboddio_b200.so (built by `python -m oddio_b200.build`, nvcc -gencode arch=compute_100a,code=sm_100a).
// ODDIO_B200_LIB_DIR overrides the default location (../../oddio_b200 relative to this crate).
use std::{env, path::PathBuf};

fn main() {
    let dir = env::var("ODDIO_B200_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../oddio_b200")
    });
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=oddio_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
    println!("cargo:rerun-if-env-changed=ODDIO_B200_LIB_DIR");
    println!("cargo:rerun-if-changed=../../include/oddio_b200.h");
}
