// build.rs — compiles the CUDA sources of rendiation_b200/csrc for sm_100a with nvcc (through the `cc` crate) and links the result,
// or — feature "prebuilt" — links the librdn_rt.so that `python -m rendiation_b200.build` left in $RDN_RT_LIB_DIR.
//
// The flags are the ones rendiation_b200/build.py uses, and they are part of the contract: -fmad=false with IEEE division / square
// root and no flush-to-zero make the device arithmetic round exactly like the reference's CPU code (DESIGN.md "Exactness");
// -ffp-contract=off does the same for the host-side builder and flattener.
use std::env;
use std::path::PathBuf;

const SOURCES: &[&str] = &[
    "bvh_builder.cpp", "accel.cpp", "traverse.cu", "compact.cu", "raygen.cu", "probe.cu", "pick.cu", "build_device.cu", "sbt.cu", "wavefront.cu", "capi.cu",
];

fn main() {
    let manifest = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    let repo = manifest.join("../..").canonicalize().expect("repository root");
    let csrc = repo.join("rendiation_b200/csrc");
    let include = repo.join("include");
    println!("cargo:rerun-if-changed={}", include.join("rdn_rt.h").display());

    if env::var_os("CARGO_FEATURE_PREBUILT").is_some() {
        let dir = env::var("RDN_RT_LIB_DIR").unwrap_or_else(|_| repo.join("rendiation_b200").display().to_string());
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-lib=dylib=rdn_rt");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
        return;
    }

    let mut build = cc::Build::new();
    build
        .cuda(true)
        .cudart("static")
        .flag("-gencode").flag("arch=compute_100a,code=sm_100a")
        .flag("-O3").flag("-lineinfo").flag("-std=c++17")
        .flag("-fmad=false").flag("-prec-div=true").flag("-prec-sqrt=true").flag("-ftz=false")
        .flag("-Xcompiler").flag("-fPIC,-ffp-contract=off,-fno-fast-math,-O2")
        .include(&include);
    for s in SOURCES {
        let p = csrc.join(s);
        println!("cargo:rerun-if-changed={}", p.display());
        build.file(p);
    }
    for h in ["accel.h", "bvh_builder.h", "kernels.h", "layout.h", "rdn_math.h"] {
        println!("cargo:rerun-if-changed={}", csrc.join(h).display());
    }
    build.compile("rdn_rt");
    println!("cargo:rustc-link-lib=dylib=stdc++");
}
