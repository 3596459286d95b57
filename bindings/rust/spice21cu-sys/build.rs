// Builds libspice21cu.so with nvcc exactly as spice21_b200/csrc/Makefile does (sm_100a, no FMA contraction) and links it.
// SPICE21CU_CSRC may point at the csrc directory; by default it is found relative to this crate inside the repository.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let here = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    let csrc = env::var("SPICE21CU_CSRC").map(PathBuf::from).unwrap_or_else(|_| here.join("../../../spice21_b200/csrc"));
    let so = out.join("libspice21cu.so");
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    let mut cmd = Command::new(nvcc);
    cmd.args(&[
        "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo", "-fmad=false",
        "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math", "-shared", "-x", "cu",
    ]);
    for f in &["capi.cpp", "kernels/newton.cu", "kernels/coop.cu", "kernels/hybrid.cu"] {
        cmd.arg(csrc.join(f));
    }
    cmd.arg("-o").arg(&so).arg("-lcudart");
    let status = cmd.status().expect("nvcc not found (set NVCC)");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=spice21cu");
    println!("cargo:rerun-if-changed={}", csrc.display());
    println!("cargo:rerun-if-env-changed=SPICE21CU_CSRC");
}
