// build.rs — what a maintainer of rust-vslam adds to build the CUDA side with nvcc for sm_100a and link it.
// (Source only: the build image of this repository has no Rust toolchain; the same nvcc line is what
//  orb-slam3-rust_b200/Makefile runs and what __graft_entry__.build() checks.)
use std::{env, path::PathBuf, process::Command};

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let csrc = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../csrc");
    let lib = out.join("libvslam_b200.so");
    let status = Command::new(env::var("NVCC").unwrap_or_else(|_| "nvcc".into()))
        .args(["-O3", "-std=c++17", "-lineinfo", "-fmad=false"])
        .args(["-gencode", "arch=compute_100a,code=sm_100a"])
        .args(["-Xcompiler", "-fPIC", "-shared", "-o"])
        .arg(&lib)
        .arg(csrc.join("vslam_b200.cu"))
        .status()
        .expect("nvcc not found: the stereo front end has no CPU fallback");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=vslam_b200");
    for f in ["vslam_b200.cu", "orb_kernels.cuh", "stereo_kernels.cuh", "track_kernels.cuh", "abi_pieces.inl",
              "select_serial.h", "vslam_params.h", "brief_pattern.inc"] {
        println!("cargo:rerun-if-changed={}", csrc.join(f).display());
    }
}
