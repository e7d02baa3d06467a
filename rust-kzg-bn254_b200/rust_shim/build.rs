// Builds libkzgbn254_b200.so with nvcc for sm_100a (csrc/Makefile) and links it.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("..");
    let status = Command::new("make")
        .arg("-C")
        .arg(root.join("csrc"))
        .arg("-j")
        .status()
        .expect("make not found");
    assert!(status.success(), "building the CUDA library failed (needs nvcc with sm_100a support)");
    println!("cargo:rustc-link-search=native={}", root.display());
    println!("cargo:rustc-link-lib=dylib=kzgbn254_b200");
    println!("cargo:rerun-if-changed=../csrc");
    println!("cargo:rerun-if-changed=../../include/kzg_bn254_b200.h");
}
