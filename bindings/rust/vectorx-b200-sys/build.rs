// Builds libvectorx_b200.so from the CUDA sources of this repository with nvcc for sm_100a (the flags of
// vectorx_b200/csrc/Makefile) and tells cargo where to find it.  VECTORX_B200_ROOT overrides the repository root;
// VECTORX_B200_LIB_DIR points at a prebuilt library instead of building.
use std::env;
use std::path::PathBuf;
use std::process::Command;

fn main() {
    println!("cargo:rerun-if-env-changed=VECTORX_B200_ROOT");
    println!("cargo:rerun-if-env-changed=VECTORX_B200_LIB_DIR");
    if let Ok(dir) = env::var("VECTORX_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-lib=dylib=vectorx_b200");
        return;
    }
    let manifest = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    let root = env::var("VECTORX_B200_ROOT").map(PathBuf::from).unwrap_or_else(|_| manifest.join("../../.."));
    let csrc = root.join("vectorx_b200/csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    let mut sources = Vec::new();
    for entry in std::fs::read_dir(&csrc).expect("vectorx_b200/csrc not found") {
        let p = entry.unwrap().path();
        if p.extension().map(|e| e == "cu").unwrap_or(false) {
            println!("cargo:rerun-if-changed={}", p.display());
            sources.push(p);
        }
    }
    let lib = out.join("libvectorx_b200.so");
    let status = Command::new(&nvcc)
        .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--use_fast_math",
               "-Xcompiler", "-fPIC,-O3", "-shared", "-o"])
        .arg(&lib)
        .args(&sources)
        .arg("-lcudart")
        .status()
        .expect("failed to run nvcc");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=vectorx_b200");
}
