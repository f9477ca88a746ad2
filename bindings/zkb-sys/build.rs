// Locates libzkb.so: $ZKB_LIB_DIR, else <repo>/ckb_zkp_b200 relative to this crate (the in-tree build of
// `make -C ckb_zkp_b200/csrc`).  The library links cudart statically; nothing else is needed at link time.
use std::{env, path::PathBuf};

fn main() {
    let dir = env::var("ZKB_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../ckb_zkp_b200")
    });
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=zkb");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
    if env::var("CARGO_FEATURE_LINK_NCCL").is_ok() {
        println!("cargo:rustc-link-lib=dylib=nccl");
    }
    println!("cargo:rerun-if-env-changed=ZKB_LIB_DIR");
    println!("cargo:rerun-if-changed=../../include/zkb.h");
}
