// Links the in-tree shared library.  WHISPER_B200_LIB_DIR overrides the default location
// (<repo>/whisper.rs_b200/csrc, where `make` leaves libwhisper_b200.so).
fn main() {
    let dir = std::env::var("WHISPER_B200_LIB_DIR").unwrap_or_else(|_| {
        let here = std::path::PathBuf::from(std::env::var("CARGO_MANIFEST_DIR").unwrap());
        here.join("../../csrc").to_string_lossy().into_owned()
    });
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=whisper_b200");
    println!("cargo:rerun-if-env-changed=WHISPER_B200_LIB_DIR");
}
