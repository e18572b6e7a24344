"""Profiling driver: <H> of the Heisenberg chain on n sites (BASELINE config 3's expectation value) -- the
batched read-only kernel k_pauli_expect_window."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quant_iron_b200 as qi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
h = qi.heisenberg_1d(n, 1.0, 2.0, 3.0, 0.5, 0.1)
st = qi.State.new_plus(n)
qi.trotter_evolve_state_(h, st, 0.01, 1, qi.TrotterOrder.First)
for _ in range(reps):
    e = h.expectation_value(st)
qi.engine.synchronize()
print("<H>", e)
