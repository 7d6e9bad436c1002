// Links libmdb200.so (built by `make -C metabodecon_rust_b200/csrc`).
fn main() {
    if let Ok(dir) = std::env::var("MDB200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    }
    println!("cargo:rustc-link-lib=dylib=mdb200");
    println!("cargo:rerun-if-env-changed=MDB200_LIB_DIR");
}
