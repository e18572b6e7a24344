"""Same-box A/B of two builds of libqiron_b200 (box-to-box variance on this pool is 3-5 %, more than most kernel
changes: compare on ONE box, alternating).

    python tools/ab_variants.py --base HEAD~1            # builds variants/libqiron_base.so (from git) and _new.so (working tree)
    python tools/ab_variants.py --base HEAD --flags "-DFOO"   # same sources, `new` compiled with extra flags

then prints the gpurun command that runs tools/exp_window.py on both, twice, alternating."""
import argparse
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "quant_iron_b200", "csrc")
OUT = os.path.join(ROOT, "quant_iron_b200", "lib", "variants")
SOURCES = ["engine", "state", "gates", "window", "pauli", "pauli_window", "measure", "shard", "host_pipeline"]
NVCC = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
        "--expt-relaxed-constexpr", "-shared"]


def build(src_dir, include_dir, name, extra):
    out = os.path.join(OUT, f"libqiron_{name}.so")
    cmd = NVCC + extra + ["-I", include_dir, "-I", src_dir, "-o", out] + [os.path.join(src_dir, s + ".cu") for s in SOURCES if os.path.exists(os.path.join(src_dir, s + ".cu"))] + ["-lcudart"]
    subprocess.check_call(cmd)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--base", default="HEAD")
    ap.add_argument("--flags", default="")
    ap.add_argument("--script", default="tools/exp_window.py")
    a = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="qiron_ab_")
    try:
        for sub in ("quant_iron_b200/csrc", "include"):
            os.makedirs(os.path.join(tmp, sub), exist_ok=True)
            files = subprocess.check_output(["git", "ls-tree", "--name-only", f"{a.base}:{sub}"], cwd=ROOT, text=True).split()
            for f in files:
                blob = subprocess.check_output(["git", "show", f"{a.base}:{sub}/{f}"], cwd=ROOT)
                with open(os.path.join(tmp, sub, f), "wb") as fh:
                    fh.write(blob)
        base = build(os.path.join(tmp, "quant_iron_b200/csrc"), os.path.join(tmp, "include"), "base", [])
        new = build(CSRC, os.path.join(ROOT, "include"), "new", a.flags.split())
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    rel = lambda p: os.path.relpath(p, ROOT)
    loop = (f"for i in 1 2; do QIRON_B200_LIB={rel(new)} timeout 100 python {a.script} --tag new | grep -v norm; "
            f"QIRON_B200_LIB={rel(base)} timeout 100 python {a.script} --tag base | grep -v norm; done")
    print("built:", rel(base), rel(new))
    print(f"/usr/local/graft/bin/gpurun --timeout 600 -- '{loop}'")
    print("(remove quant_iron_b200/lib/variants/ afterwards: it travels to the GPU box with every call)")


if __name__ == "__main__":
    sys.exit(main())
