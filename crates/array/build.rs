//! Builds libagpu.so (the CUDA side of the array crate) with nvcc for sm_100a and links it.
//! Replaces nothing in the reference — its array crate had no build script because naga compiled
//! the WGSL at run time (crates/array/src/gpu_utils/gpu_device.rs:137-168).
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let csrc = root.join("arrow_gpu_b200/csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    let sources = ["device", "arith", "unary", "compare", "logical", "cast", "routines", "reduce", "chain", "chain_int", "exchange"];
    let mut objects = Vec::new();
    for s in sources {
        let obj = out.join(format!("{s}.o"));
        let status = Command::new(&nvcc)
            .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-c"])
            .arg(csrc.join(format!("{s}.cu")))
            .arg("-o")
            .arg(&obj)
            .status()
            .expect("nvcc not found: the array crate needs the CUDA toolkit (there is no CPU fallback)");
        assert!(status.success(), "nvcc failed on {s}.cu");
        println!("cargo:rerun-if-changed={}", csrc.join(format!("{s}.cu")).display());
        objects.push(obj);
    }
    for h in ["common.cuh", "elementwise.cuh", "bits.cuh", "ops.cuh"] {
        println!("cargo:rerun-if-changed={}", csrc.join(h).display());
    }
    println!("cargo:rerun-if-changed={}", root.join("include/agpu.h").display());
    let lib = out.join("libagpu.so");
    let status = Command::new(&nvcc)
        .args(["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static", "-o"])
        .arg(&lib)
        .args(&objects)
        .status()
        .unwrap();
    assert!(status.success(), "linking libagpu.so failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=agpu");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", out.display());
}
