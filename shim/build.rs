// build.rs for the reference crate with the CUDA backend enabled (cargo feature `cuda`).
// Compiles the B200 library from this repository's sources with nvcc for sm_100a and links it.
// COMPILE-UNVERIFIED (no cargo in the image this was written in); the same nvcc command line is what
// rust-pathtracer_b200/csrc/Makefile runs and what the driver's build check exercises.
//
//   RPT_B200_DIR   path to a checkout of this repository (default: ../b200-spectral-pt)
//   RPT_B200_LIB   use a prebuilt librpt_b200.so in that directory instead of invoking nvcc
//   NVCC           nvcc binary (default: /usr/local/cuda/bin/nvcc)
use std::env;
use std::path::PathBuf;
use std::process::Command;

fn main() {
    if env::var("CARGO_FEATURE_CUDA").is_err() {
        return;
    }
    let root = PathBuf::from(env::var("RPT_B200_DIR").unwrap_or_else(|_| "../b200-spectral-pt".into()));
    let csrc = root.join("rust-pathtracer_b200").join("csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    println!("cargo:rerun-if-env-changed=RPT_B200_DIR");
    println!("cargo:rerun-if-env-changed=RPT_B200_LIB");
    for f in ["rpt_kernels.cu", "rpt_multi.cu", "rpt_device.cuh", "rpt_bvh.cpp", "rpt_bvh.h"] {
        println!("cargo:rerun-if-changed={}", csrc.join(f).display());
    }
    println!("cargo:rerun-if-changed={}", root.join("include").join("rpt.h").display());
    let lib_dir = if let Ok(prebuilt) = env::var("RPT_B200_LIB") {
        PathBuf::from(prebuilt)
    } else {
        let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
        let status = Command::new(&nvcc)
            .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
                   "-fmad=false", // Rust never contracts a*b+c: the device arithmetic must not either (DESIGN.md "Numerics")
                   "-Xcompiler", "-fPIC", "-shared", "-o"])
            .arg(out.join("librpt_b200.so"))
            .arg(csrc.join("rpt_kernels.cu"))
            .arg(csrc.join("rpt_multi.cu"))
            .arg(csrc.join("rpt_bvh.cpp"))
            .arg("-ldl")
            .status()
            .expect("failed to run nvcc (set NVCC, or RPT_B200_LIB to a directory with a prebuilt librpt_b200.so)");
        assert!(status.success(), "nvcc failed");
        out.clone()
    };
    println!("cargo:rustc-link-search=native={}", lib_dir.display());
    println!("cargo:rustc-link-lib=dylib=rpt_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", lib_dir.display());
    // CUDA runtime: nvcc links cudart statically into the .so by default; NCCL is dlopen'ed by the library on first
    // multi-GPU use (libnccl.so.2 or $RPT_NCCL_LIB), so neither appears here.
}
