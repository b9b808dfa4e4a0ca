// build.rs of the Rust shim (source only: there is no Rust toolchain in the environment this repository was built in).
// Builds libbp_b200.so from the CUDA sources with the repository's own Makefile (nvcc for sm_100a) and links it.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let csrc = root.join("ark_bulletproofs_b200/csrc");
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    let status = Command::new("make").arg("-C").arg(&csrc).arg("-j8").env("NVCC", nvcc).status().expect("make");
    assert!(status.success(), "building libbp_b200.so failed");
    println!("cargo:rustc-link-search=native={}", root.join("ark_bulletproofs_b200").display());
    println!("cargo:rustc-link-lib=dylib=bp_b200");
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rerun-if-changed={}", csrc.display());
    println!("cargo:rerun-if-changed={}", root.join("include/bp_b200.h").display());
}
