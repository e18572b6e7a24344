"""Host-only: passes the Pauli-exp batching scheduler builds for a Heisenberg chain (no GPU needed)."""
import argparse
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import quant_iron_b200 as qi  # noqa: E402
from quant_iron_b200 import _ffi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--sites", type=int, default=24)
ap.add_argument("--steps", type=int, default=50)
a = ap.parse_args()
h = qi.heisenberg_1d(a.sites, 1.0, 2.0, 3.0, 0.5, 0.1)
arr, n, keep = h.term_array()
rows = (C.c_int32 * 100000)()
cnt = C.c_uint64(0)
t0 = time.time()
_ffi.check(_ffi.lib.qi_debug_pauli_schedule(a.sites, arr, n, a.steps, rows, 100000, C.byref(cnt)))
dt = time.time() - t0
per = [rows[i] for i in range(cnt.value)]
print(f"{n} terms x {a.steps} steps = {n * a.steps} exps -> {len(per)} passes ({len(per) / a.steps:.2f} per step), host {dt * 1e3:.1f} ms")
print("terms per pass (first 40):", per[:40])
