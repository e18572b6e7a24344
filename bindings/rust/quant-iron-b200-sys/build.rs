// Tells rustc where libqiron_b200.so lives: QIRON_B200_LIB_DIR, or the in-tree build directory.
fn main() {
    let dir = std::env::var("QIRON_B200_LIB_DIR").unwrap_or_else(|_| {
        let here = std::env::var("CARGO_MANIFEST_DIR").unwrap();
        format!("{}/../../../quant_iron_b200/lib", here)
    });
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=qiron_b200");
    println!("cargo:rerun-if-env-changed=QIRON_B200_LIB_DIR");
}
