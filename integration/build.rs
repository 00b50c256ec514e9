// needle/build.rs -- builds libneedle_b200.a from the CUDA sources with nvcc and links it.
//
// Replaces nothing in the reference (needle has no build script of its own; needle-capi's
// build.rs, needle-capi/build.rs:1-17, only runs cbindgen and is untouched).  Enabled by the
// cargo feature `b200` (integration/Cargo.toml.patch).
//
// NCCL is NOT linked: the multi-GPU jobs (nb200_comm_* / nb200_mjob_*) load libnccl.so.2 with
// dlopen at first use, so a single-GPU machine without NCCL still runs everything else.
use std::path::PathBuf;
use std::process::Command;

const SOURCES: &[&str] = &[
    "api.cu", "match.cu", "fingerprint.cu", "vote_device.cu", "multi.cu", "vote.cpp", "persist.cpp",
];

fn main() {
    if std::env::var_os("CARGO_FEATURE_B200").is_none() {
        return;
    }
    // where the needle-b200 checkout lives (default: vendored next to the crate)
    let root = PathBuf::from(std::env::var("NEEDLE_B200_DIR").unwrap_or_else(|_| "../needle-b200".into()));
    let csrc = root.join("needle_b200").join("csrc");
    let include = root.join("include");
    let out = PathBuf::from(std::env::var("OUT_DIR").unwrap());
    let nvcc = std::env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    let cuda_lib = std::env::var("CUDA_LIB_DIR").unwrap_or_else(|_| "/usr/local/cuda/lib64".into());

    let mut objects = Vec::new();
    for s in SOURCES {
        let src = csrc.join(s);
        let obj = out.join(s).with_extension("o");
        println!("cargo:rerun-if-changed={}", src.display());
        let status = Command::new(&nvcc)
            .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "--fmad=true"])
            .args(["-Xcompiler", "-fPIC,-fno-fast-math,-ffp-contract=off,-mpopcnt"])
            .arg("-I").arg(&include)
            .arg("-c").arg(&src)
            .arg("-o").arg(&obj)
            .status()
            .expect("nvcc not found (set NVCC)");
        assert!(status.success(), "nvcc failed on {}", src.display());
        objects.push(obj);
    }
    for h in ["common.h", "fp_tables.h", "fp_chroma_fold.inc"] {
        println!("cargo:rerun-if-changed={}", csrc.join(h).display());
    }
    println!("cargo:rerun-if-changed={}", include.join("needle_b200.h").display());

    let lib = out.join("libneedle_b200.a");
    let _ = std::fs::remove_file(&lib);
    let status = Command::new("ar").arg("rcs").arg(&lib).args(&objects).status().expect("ar not found");
    assert!(status.success());

    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=needle_b200");
    println!("cargo:rustc-link-search=native={}", cuda_lib);
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rustc-link-lib=dylib=stdc++");
    println!("cargo:rustc-link-lib=dylib=dl");      // dlopen("libnccl.so.2")
    println!("cargo:rustc-link-lib=dylib=pthread");
}
