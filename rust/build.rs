//! Builds libbn254v.a with nvcc (sm_100a) from this repository's CUDA sources, one translation unit per kernel family,
//! and links it together with the CUDA runtime.  With `--features prebuilt`, links the shared library found in
//! BN254V_LIB_DIR instead (what `python snark-bn254-verifier_b200/build.py` produces).
use std::{env, path::PathBuf, process::Command};

const SOURCES: [&str; 6] = ["bn254v.cu", "k_groth16.cu", "k_groth16_agg.cu", "k_plonk.cu", "k_pairing.cu", "k_aux.cu"];

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("..");
    let csrc = root.join("snark-bn254-verifier_b200").join("csrc");
    println!("cargo:rerun-if-changed={}", csrc.display());
    println!("cargo:rerun-if-changed={}", root.join("include").display());
    println!("cargo:rerun-if-env-changed=BN254V_LIB_DIR");
    if env::var("CARGO_FEATURE_PREBUILT").is_ok() {
        let dir = env::var("BN254V_LIB_DIR").unwrap_or_else(|_| root.join("snark-bn254-verifier_b200").display().to_string());
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-lib=dylib=bn254v");
        return;
    }
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    let mut objects = Vec::new();
    let mut children = Vec::new();
    for src in SOURCES {
        let obj = out.join(src.replace(".cu", ".o"));
        let child = Command::new(&nvcc)
            .args(["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-c", "-o"])
            .arg(&obj)
            .arg(csrc.join(src))
            .spawn()
            .expect("nvcc not found (set NVCC or use --features prebuilt)");
        children.push(child);
        objects.push(obj);
    }
    for mut c in children {
        assert!(c.wait().expect("nvcc").success(), "nvcc failed");
    }
    let lib = out.join("libbn254v.a");
    let status = Command::new(&nvcc).args(["-lib", "-o"]).arg(&lib).args(&objects).status().expect("nvcc -lib");
    assert!(status.success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=bn254v");
    if let Ok(cuda) = env::var("CUDA_HOME") {
        println!("cargo:rustc-link-search=native={cuda}/lib64");
    } else {
        println!("cargo:rustc-link-search=native=/usr/local/cuda/lib64");
    }
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rustc-link-lib=dylib=stdc++");
}
