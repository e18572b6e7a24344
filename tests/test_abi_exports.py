"""CPU-side checks of the C-ABI boundary: the library loads, exports every symbol the header
declares, and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "qiron_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(qi_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from quant_iron_b200 import _ffi
    lib = ctypes.CDLL(_ffi.LIB_PATH)
    names = _declared()
    assert len(names) >= 50
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_header_cites_reference_for_each_entry_group():
    src = open(HEADER).read()
    for needle in ("operator.rs:214-273", "state.rs:525-730", "pauli_string.rs:237-262",
                   "circuit.rs:160-172", "errors.rs:3-97", "time_evolution.rs:140-167"):
        assert needle in src


def test_no_cpu_fallback_without_device():
    import quant_iron_b200 as qi
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(qi.Error) as e:
        qi.State.new_zero(3)
    assert e.value.variant == "CudaError"


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "quant_iron_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "libqi_oracle" not in text and "qi_oracle.c" not in text and "orc_" not in text, f


def test_host_side_validation_without_device():
    """Errors raised by the host mirror before any device call."""
    import quant_iron_b200 as qi
    with pytest.raises(qi.Error) as e:
        qi.heisenberg_1d(1, 1.0, 2.0, 3.0, 4.0, 5.0)
    assert (e.value.variant, e.value.payload) == ("InvalidNumberOfInputs", (1, 2))
    with pytest.raises(qi.Error) as e:
        qi.Circuit.with_gates([qi.Gate.h_gate(1), qi.Gate.cnot_gate(0, 3)], 2)
    assert (e.value.variant, e.value.payload) == ("InvalidQubitIndex", (3, 2))
    with pytest.raises(qi.Error) as e:
        qi.Unitary2.new([[0j, 1 + 0j], [1 + 0j, 1 + 0j]])
    assert e.value.variant == "NonUnitaryMatrix"
    sub = qi.Subroutine.qft(list(range(20)), 20)
    assert len(sub.gates) == 220          # 20 H + 190 CP + 10 SWAP
    assert len(qi.heisenberg_1d(24, 1.0, 2.0, 3.0, 0.5, 0.1).terms) == 96
    specs = qi.workloads.random_layered_circuit(28, 40)
    assert len(specs) == 40 * 28 + 20 * 14 + 20 * 13


def test_rust_sys_bindings_match_the_header():
    """the Rust `-sys` crate source is generated from include/qiron_b200.h; the committed files must
    be what the generator produces today and must declare every exported function."""
    import re
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "gen_rust_sys.py"), "--check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    lib_rs = open(os.path.join(root, "bindings", "rust", "quant-iron-b200-sys", "src", "lib.rs")).read()
    header = open(os.path.join(root, "include", "qiron_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", " ", header, flags=re.S)
    names = set(re.findall(r"\b(qi_[a-z0-9_]+)\s*\(", header))
    assert len(names) >= 59
    for n in names:
        assert re.search(rf"pub fn {n}\(", lib_rs), n


def test_struct_layouts_agree_between_c_rust_asserts_and_ctypes():
    """The generated crate asserts size / alignment / field offsets of every C struct at compile time (gcc's layout of the header
    on x86_64); the same numbers must hold for the ctypes mirror this repository's Python host layer uses."""
    import ctypes as C
    import re

    from quant_iron_b200 import _ffi
    src = open(os.path.join(ROOT, "bindings", "rust", "quant-iron-b200-sys", "src", "lib.rs")).read()
    sizes = {m.group(1): int(m.group(2)) for m in re.finditer(r"size_of::<(\w+)>\(\) == (\d+)", src)}
    offsets = {(m.group(1), m.group(2)): int(m.group(3)) for m in re.finditer(r"offset_of!\((\w+), (\w+)\) == (\d+)", src)}
    assert {"qi_gate", "qi_pauli_term", "qi_kernel_stat"} <= set(sizes)
    mirrors = {"qi_gate": _ffi.QiGate, "qi_pauli_term": _ffi.QiPauliTerm, "qi_kernel_stat": _ffi.QiKernelStat}
    for name, cls in mirrors.items():
        assert C.sizeof(cls) == sizes[name], name
        for fname, _ in cls._fields_:
            assert getattr(cls, fname).offset == offsets[(name, fname)], (name, fname)
