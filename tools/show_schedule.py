"""Host-only: print how the fused executor splits a workload into passes (no GPU needed)."""
import argparse
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quant_iron_b200 as qi  # noqa: E402
from quant_iron_b200 import _ffi, workloads as w  # noqa: E402


def schedule(n, specs, regs=4):
    c = w.build_circuit(qi, n, specs)
    recs = [g.op.record(g.targets, g.controls) for g in c.gates]
    arr = (_ffi.QiGate * len(recs))()
    for i, (r, _k) in enumerate(recs):
        arr[i] = r
    rows = (C.c_int32 * (8 * 4096))()
    nrows = C.c_uint64()
    _ffi.check(_ffi.lib.qi_debug_schedule(n, arr, len(recs), regs, rows, 4096, C.byref(nrows)))
    return [tuple(rows[8 * i + k] for k in range(7)) for i in range(min(4096, nrows.value))]


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--qubits", type=int, default=30)
    ap.add_argument("--what", default="layered")
    ap.add_argument("--regs", type=int, default=4)
    a = ap.parse_args()
    specs = w.random_layered_circuit(a.qubits, 40) if a.what == "layered" else w.qft_specs(a.qubits)
    rows = schedule(a.qubits, specs, a.regs)
    print(f"{len(specs)} gates -> {len(rows)} passes")
    print("simple regs lane reg diag table absorbed_cnots")
    for r in rows:
        print(*r)
