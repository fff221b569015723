// build.rs -- compiles the CUDA sources of tiny-ram-halo2_b200/csrc for sm_100a with nvcc and links libtrp.
// Mirrors tiny-ram-halo2_b200/csrc/Makefile.  UNTESTED in the build image (no rustc); kept compile-ready.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("..");
    let csrc = root.join("tiny-ram-halo2_b200").join("csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    let mut objs = Vec::new();
    for unit in ["capi", "ntt", "msm", "quotient", "microbench"] {
        let src = csrc.join(format!("{unit}.cu"));
        let obj = out.join(format!("{unit}.o"));
        println!("cargo:rerun-if-changed={}", src.display());
        let ok = Command::new(&nvcc)
            .args(["-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
                   "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-c"])
            .arg(&src).arg("-o").arg(&obj)
            .status().expect("nvcc not found").success();
        assert!(ok, "nvcc failed on {}", src.display());
        objs.push(obj);
    }
    let lib = out.join("libtrp.so");
    let ok = Command::new(&nvcc)
        .args(["-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o"]).arg(&lib).args(&objs).arg("-lcudart")
        .status().expect("nvcc not found").success();
    assert!(ok, "linking libtrp.so failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=trp");
    println!("cargo:rustc-link-search=native=/usr/local/cuda/lib64");
    println!("cargo:rustc-link-lib=dylib=cudart");
}
