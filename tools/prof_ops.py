"""ncu driver: ONE window pass holding m identical gates (see tools/exp_opcost.py)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quant_iron_b200 as qi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--qubits", type=int, default=30)
ap.add_argument("--kind", default="h_reg")
ap.add_argument("--m", type=int, default=32)
a = ap.parse_args()
n = a.qubits
st = qi.State.new_random(n)
qs = [0, 1, 2, 3] if a.kind.endswith("lane") else [10, 14, 19, 25]
b = qi.CircuitBuilder(n)
for i in range(a.m):
    q = qs[i % 4]
    if a.kind.startswith("h"):
        b.h_gate(q)
    elif a.kind.startswith("rx"):
        b.rx_gate(q, 0.1 + 0.01 * i)
    else:
        b.cnot_gate(q, 7 + (i % 3))
c = b.build()
c.execute_(st)
c.execute_(st)
qi.engine.synchronize()
print("norm", st.norm_sqr())
