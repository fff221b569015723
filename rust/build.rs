// build.rs -- compiles the CUDA sources of tiny-ram-halo2_b200/csrc for sm_100a with nvcc and links libtrp.
// The unit list is READ from csrc/Makefile's `OBJS :=` line, so the crate links exactly what the Makefile links
// (tests/test_rust_shim_cpu.py checks that every listed unit exists).  Not run in the build image (no rustc).
use std::{env, fs, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("..");
    let csrc = root.join("tiny-ram-halo2_b200").join("csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    let makefile = fs::read_to_string(csrc.join("Makefile")).expect("csrc/Makefile");
    println!("cargo:rerun-if-changed={}", csrc.join("Makefile").display());
    let units: Vec<String> = makefile
        .lines()
        .find(|l| l.starts_with("OBJS :="))
        .expect("csrc/Makefile has no OBJS := line")
        .trim_start_matches("OBJS :=")
        .split_whitespace()
        .map(|o| o.trim_end_matches(".o").to_string())
        .collect();
    assert!(!units.is_empty());
    for entry in fs::read_dir(&csrc).unwrap() {
        let p = entry.unwrap().path();
        if p.extension().map_or(false, |e| e == "cuh" || e == "h") { println!("cargo:rerun-if-changed={}", p.display()); }
    }
    println!("cargo:rerun-if-changed={}", root.join("include").join("tr_prover.h").display());
    let mut objs = Vec::new();
    for unit in &units {
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
