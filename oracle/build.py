"""Build recipe for the CPU oracle (test infrastructure, never shipped).

`oracle/_build/libqi_oracle.so` is compiled from oracle/qi_oracle.c with gcc.
The reference itself (Rust, /root/reference) cannot be compiled in this image
(no cargo/rustc), so there is no oracle/_ref/ artefact: the oracle is a "port",
pinned against the reference's own known-answer tests (tests/test_ref_ported_*).
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libqi_oracle.so")
SRC = os.path.join(HERE, "qi_oracle.c")


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if (not force and os.path.exists(LIB)
            and os.path.getmtime(LIB) >= os.path.getmtime(SRC)):
        return LIB
    # -ffp-contract=off: Rust never contracts a*b+c into an FMA; neither may the oracle.
    cmd = ["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off",
           "-fno-fast-math", "-o", LIB, SRC, "-lm"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
