// Builds libpsim_b200.so from the CUDA sources with nvcc for sm_100a and links it.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let csrc = root.join("particlesim_b200/csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    println!("cargo:rerun-if-changed={}", csrc.display());
    println!("cargo:rerun-if-changed={}", root.join("include/psim_b200.h").display());
    if env::var("CARGO_FEATURE_PREBUILT").is_ok() {
        println!("cargo:rustc-link-search=native={}", root.join("particlesim_b200").display());
    } else {
        let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
        let status = Command::new(nvcc)
            .args(["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17"])
            .args(["-Xcompiler", "-fPIC", "-shared", "-o"])
            .arg(out.join("libpsim_b200.so"))
            .arg(csrc.join("api.cu"))
            .status()
            .expect("nvcc not found: the B200 force path needs the CUDA toolkit");
        assert!(status.success(), "nvcc failed");
        println!("cargo:rustc-link-search=native={}", out.display());
    }
    println!("cargo:rustc-link-lib=dylib=psim_b200");
    println!("cargo:rustc-link-lib=dylib=cudart");
}
