// Links libb200zk.so; B200ZK_LIB_DIR = directory holding it (halo2-experiments_b200/ of the b200zk repository).
fn main() {
    if let Ok(dir) = std::env::var("B200ZK_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    }
    println!("cargo:rustc-link-lib=dylib=b200zk");
    println!("cargo:rerun-if-env-changed=B200ZK_LIB_DIR");
}
