// build.rs — compiles the .cu sources with nvcc for sm_100a and links the result.
// SOURCE ONLY (no Rust toolchain in the build image); mirrors dxt_lossless_transform_b200/build.py.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let manifest = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    let csrc = manifest.join("../../dxt_lossless_transform_b200/csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let lib = out.join("libdxt_lossless_transform_cuda.so");
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    let sources = ["bcn_kernels.cu", "host_pipeline.cu", "estimator.cu", "auto_search.cu", "cabi.cu", "file_formats.cu"];

    let status = Command::new(&nvcc)
        .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17"])
        .args(["-Xcompiler", "-fPIC,-fvisibility=hidden,-O3", "--threads", "0", "-shared", "-o"])
        .arg(&lib)
        .args(sources.iter().map(|s| csrc.join(s)))
        .status()
        .expect("nvcc not found: set NVCC or put CUDA 12.9+ on PATH (there is no CPU fallback)");
    assert!(status.success(), "nvcc failed");

    for s in sources {
        println!("cargo:rerun-if-changed={}", csrc.join(s).display());
    }
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=dxt_lossless_transform_cuda");
}
