// build.rs — compile the hand-written sm_100a CUDA and link it (no CPU fallback: with the `gpu` feature on, the build
// fails without nvcc).  Sources: rasterize_b200/csrc (copied or vendored next to Cargo.toml).
use std::{env, path::PathBuf, process::Command};

const SOURCES: &[&str] = &["flatten.cu", "scan.cu", "raster.cu", "scene.cu", "small.cu", "compose.cu", "compact.cu", "stroke.cu", "parse.cu", "context.cu", "host_simd.cpp"];

fn main() {
    if env::var_os("CARGO_FEATURE_GPU").is_none() {
        return;
    }
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let csrc = PathBuf::from("rasterize_b200/csrc");
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    let mut objs = Vec::new();
    for f in SOURCES {
        let obj = out.join(format!("{f}.o"));
        let ok = Command::new(&nvcc)
            .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "--expt-relaxed-constexpr",
                   "-Xcompiler", "-fPIC", "-c"])
            // stroke.cu / parse.cu restate f64 expressions of the reference with plain operators: no multiply-add contraction there
            .args(if *f == "stroke.cu" || *f == "parse.cu" { &["--fmad=false"][..] } else { &[][..] })
            .arg(csrc.join(f))
            .arg("-o")
            .arg(&obj)
            .status()
            .expect("nvcc not found (the gpu feature has no CPU fallback)")
            .success();
        assert!(ok, "nvcc failed on {f}");
        objs.push(obj);
    }
    let lib = out.join("librasterize_b200.a");
    assert!(Command::new("ar").arg("crs").arg(&lib).args(&objs).status().unwrap().success());
    let cuda = env::var("CUDA_HOME").unwrap_or_else(|_| "/usr/local/cuda".into());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=rasterize_b200");
    println!("cargo:rustc-link-search=native={cuda}/lib64");
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rustc-link-lib=dylib=stdc++");
    println!("cargo:rerun-if-changed=rasterize_b200/csrc");
    println!("cargo:rerun-if-changed=include/rasterize_b200.h");
}
