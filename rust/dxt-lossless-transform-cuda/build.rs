// build.rs — compiles the .cu sources with nvcc for sm_100a and links the result.
// SOURCE ONLY (no Rust toolchain in the build image).  The list of translation units and the nvcc flags are READ from
// csrc/SOURCES.txt and csrc/NVCC_FLAGS.txt, the files dxt_lossless_transform_b200/build.py reads, so this build and
// the Python-driven one cannot drift apart (round 1 shipped a stale hand-written copy of the list here).
use std::{env, fs, path::PathBuf, process::Command};

fn lines(path: &PathBuf) -> Vec<String> {
    fs::read_to_string(path)
        .unwrap_or_else(|e| panic!("{}: {e}", path.display()))
        .lines()
        .map(str::trim)
        .filter(|l| !l.is_empty() && !l.starts_with('#'))
        .map(String::from)
        .collect()
}

fn main() {
    let manifest = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    let csrc = manifest.join("../../dxt_lossless_transform_b200/csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let lib = out.join("libdxt_lossless_transform_cuda.so");
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    let sources = lines(&csrc.join("SOURCES.txt"));
    let flags = lines(&csrc.join("NVCC_FLAGS.txt"));

    let status = Command::new(&nvcc)
        .args(&flags)
        .args(["-shared", "-o"])
        .arg(&lib)
        .args(sources.iter().map(|s| csrc.join(s)))
        .status()
        .expect("nvcc not found: set NVCC or put CUDA 12.9+ on PATH (there is no CPU fallback)");
    assert!(status.success(), "nvcc failed");

    for s in sources.iter().map(String::as_str).chain(["SOURCES.txt", "NVCC_FLAGS.txt"]) {
        println!("cargo:rerun-if-changed={}", csrc.join(s).display());
    }
    for h in fs::read_dir(&csrc).unwrap().flatten() {
        if h.path().extension().map_or(false, |e| e == "h") {
            println!("cargo:rerun-if-changed={}", h.path().display());
        }
    }
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=dxt_lossless_transform_cuda");
    println!("cargo:rustc-link-lib=dylib=dl");
}
