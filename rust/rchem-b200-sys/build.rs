// build.rs -- link librchem_b200.so.  Plays the role of the reference's build.rs:9-11 (cmake +
// `cargo:rustc-link-lib=static=pyquante2`): the native library is built by
// `make -C rchem_b200/csrc` (nvcc, sm_100a), this script only tells cargo where it is.
//
//   RCHEM_B200_LIB_DIR   directory holding librchem_b200.so (default: ../../rchem_b200)
//
// The bindings are hand-written in src/lib.rs from include/rchem_eri.h (no bindgen needed: the
// header is ~40 plain-C prototypes), so this crate has no build-dependencies.
use std::env;
use std::path::PathBuf;

fn main() {
    let manifest = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    let lib_dir = env::var("RCHEM_B200_LIB_DIR")
        .map(PathBuf::from)
        .unwrap_or_else(|_| manifest.join("..").join("..").join("rchem_b200"));
    println!("cargo:rustc-link-search=native={}", lib_dir.display());
    println!("cargo:rustc-link-lib=dylib=rchem_b200");
    // so `cargo test` / `cargo run` find the shared object without LD_LIBRARY_PATH
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", lib_dir.display());
    println!("cargo:rerun-if-env-changed=RCHEM_B200_LIB_DIR");
    println!("cargo:rerun-if-changed=../../include/rchem_eri.h");
}
