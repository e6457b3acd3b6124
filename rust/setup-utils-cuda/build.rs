// Builds libsnarksetup_b200.so for sm_100a with nvcc (the Makefile of the snark-setup_b200 checkout:
// `-gencode arch=compute_100a,code=sm_100a`) and links it.  The reference's own build scripts only detect the
// rustc channel (phase1-cli/build.rs:1-12); this is the "thin C-ABI/FFI layer built from build.rs".
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(
        env::var("SNARK_SETUP_B200_DIR").expect("set SNARK_SETUP_B200_DIR to the snark-setup_b200 checkout (holds include/ and snark-setup_b200/csrc/)"),
    );
    let csrc = root.join("snark-setup_b200").join("csrc");
    if env::var("SNARK_SETUP_B200_PREBUILT").is_err() {
        let status = Command::new("make")
            .arg("-j8")
            .arg("-C")
            .arg(&csrc)
            .status()
            .expect("could not run make for libsnarksetup_b200.so");
        assert!(status.success(), "nvcc build of libsnarksetup_b200.so failed (needs CUDA >= 12.8 for sm_100a)");
    }
    println!("cargo:rustc-link-search=native={}", csrc.display());
    println!("cargo:rustc-link-lib=dylib=snarksetup_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", csrc.display());
    println!("cargo:rerun-if-env-changed=SNARK_SETUP_B200_DIR");
    println!("cargo:rerun-if-changed={}", csrc.display());
    println!("cargo:rerun-if-changed={}", root.join("include").join("snark_setup_b200.h").display());
}
