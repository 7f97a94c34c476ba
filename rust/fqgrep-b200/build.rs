// Builds libfqgrep_b200.so with nvcc for sm_100a (through the repo's own build script, which holds the file list and the
// flags: -gencode arch=compute_100a,code=sm_100a -lineinfo -O3) and links it.  Set FQGREP_B200_REPO to the checkout when
// the crate is vendored outside of it, NVCC to pick a compiler, FQGREP_B200_SKIP_BUILD=1 to link a prebuilt library.
// NOTE: the build image has no cargo/rustc, so this script has never run.
use std::path::PathBuf;
use std::process::Command;

fn main() {
    let repo = std::env::var("FQGREP_B200_REPO")
        .map(PathBuf::from)
        .unwrap_or_else(|_| PathBuf::from(env!("CARGO_MANIFEST_DIR")).join("../.."));
    let lib_dir = repo.join("fqgrep_b200");
    if std::env::var("FQGREP_B200_SKIP_BUILD").is_err() {
        let status = Command::new(std::env::var("PYTHON").unwrap_or_else(|_| "python3".into()))
            .current_dir(&repo)
            .args(["-m", "fqgrep_b200.build"])
            .status()
            .expect("could not run `python3 -m fqgrep_b200.build` (needs nvcc with sm_100a support)");
        assert!(status.success(), "nvcc build of libfqgrep_b200.so failed");
    }
    println!("cargo:rustc-link-search=native={}", lib_dir.display());
    println!("cargo:rustc-link-lib=dylib=fqgrep_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", lib_dir.display());
    for f in ["api.cu", "fastq_pipeline.cu", "stream.cu", "match_soa.cu", "match_fastq.cu", "regex_compile.cc", "ac_build.cc", "fastq_cut.cc"] {
        println!("cargo:rerun-if-changed={}", lib_dir.join("csrc").join(f).display());
    }
    println!("cargo:rerun-if-changed={}", repo.join("include/fqgrep_b200.h").display());
    println!("cargo:rerun-if-env-changed=FQGREP_B200_REPO");
}
