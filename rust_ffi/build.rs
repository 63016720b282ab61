// Links libising_b200.so (built by `python __graft_entry__.py build` with nvcc for sm_100a).
fn main() {
    let dir = std::env::var("ISING_B200_LIB_DIR")
        .unwrap_or_else(|_| "../pyisingmontecarlo_b200".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=ising_b200");
    println!("cargo:rerun-if-env-changed=ISING_B200_LIB_DIR");
}
