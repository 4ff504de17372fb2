// Links the prebuilt CUDA library (built with `make -C mpc-jellyfish_b200/csrc`; nvcc, sm_100a).
// JF_B200_LIB_DIR points at the directory holding libjf_b200.so.
fn main() {
    let dir = std::env::var("JF_B200_LIB_DIR").expect("set JF_B200_LIB_DIR to the directory of libjf_b200.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=jf_b200");
    println!("cargo:rerun-if-env-changed=JF_B200_LIB_DIR");
    #[cfg(feature = "regenerate")]
    {
        let header = std::env::var("JF_B200_HEADER").unwrap_or_else(|_| "../../include/jf_b200.h".into());
        bindgen::Builder::default()
            .header(header)
            .allowlist_function("jf_.*")
            .allowlist_type("jf_.*")
            .generate()
            .expect("bindgen")
            .write_to_file("src/bindings.rs")
            .expect("write bindings");
    }
}
