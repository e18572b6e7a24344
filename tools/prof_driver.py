"""Short driver for ncu captures: a few passes of each kernel family at N qubits (default 30).
Usage (under gpurun): ncu ... python tools/prof_driver.py [--qubits 30] [--what window|single|all]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quant_iron_b200 as qi  # noqa: E402
from quant_iron_b200 import workloads as w  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--qubits", type=int, default=30)
ap.add_argument("--what", default="all")
ap.add_argument("--depth", type=int, default=4)
a = ap.parse_args()
n = a.qubits
st = qi.State.new_random(n)
if a.what in ("window", "all"):
    c = w.build_circuit(qi, n, w.random_layered_circuit(n, a.depth))
    c.execute_(st)
    c.execute_(st)
if a.what in ("single", "all"):
    for t in (0, 3, 5, 12, n - 1):
        st.h_(t)
        st.rx_(t, 0.3)
        st.rz_(t, 0.3)
    st.cnot_(n - 1, 0)
    st.cp_multi_([3], [n - 2], 0.3)
if a.what in ("simple", "all"):
    qi.engine.set_option("path", 1)
    for t in (0, 3, 5, 12, n - 1):
        st.h_(t)
        st.rz_(t, 0.3)
    st.cnot_(n - 1, 0)
    qi.engine.set_option("path", 0)
if a.what in ("qft", "all"):
    qft = qi.CircuitBuilder(n).add_subroutine(qi.Subroutine.qft(list(range(n)), n)).build()
    qft.execute_(st)
qi.engine.synchronize()
print("norm", st.norm_sqr(), qi.engine.stats())
