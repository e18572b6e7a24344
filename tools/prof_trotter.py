"""Profiling driver: a few first-order Trotter steps of the 24-site Heisenberg chain (BASELINE config 3)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quant_iron_b200 as qi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
h = qi.heisenberg_1d(n, 1.0, 2.0, 3.0, 0.5, 0.1)
st = qi.State.new_plus(n)
qi.trotter_evolve_state_(h, st, 0.01, steps, qi.TrotterOrder.First)
qi.engine.synchronize()
print("norm", st.norm_sqr())
