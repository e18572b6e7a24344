"""Measurement path at n qubits: marginal probabilities over a qubit set and a collapse, CUDA-event times and GB/s.  python tools/prof_measure.py [n]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quant_iron_b200 as qi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
qi.engine.init(0)
st = qi.State.new_random(n)
for qubits in ([0, 7, 13, 22, n - 1], [7, 13, 22, n - 1], [0, 1, 2, 13, n - 1], [5, 6]):
    st.probabilities(qubits)
    qi.engine.timer_start()
    for _ in range(5):
        p = st.probabilities(qubits)
    ms = qi.engine.timer_stop() / 5
    print(f"n={n} probabilities{qubits}: {ms:.2f} ms = {16.0 * (1 << n) / ms / 1e6:.0f} GB/s, sum-1={abs(p.sum() - 1):.1e}", flush=True)
for qubits in ([0, 7, 13, 22, n - 1], [7, 13]):
    s2 = qi.State.new_random(n)
    qi.engine.synchronize()
    qi.engine.timer_start()
    s2.measure_(qi.MeasurementBasis.Computational, qubits, seed=3)
    ms = qi.engine.timer_stop()
    print(f"n={n} measure_{qubits} (probabilities + scan + sample + collapse + normalise): {ms:.2f} ms, norm={s2.norm_sqr():.15f}", flush=True)
    del s2
