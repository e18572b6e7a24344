"""30-qubit single-gate pass timings (GB/s = algorithmic bytes / CUDA-event time)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quant_iron_b200 as qi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--qubits", type=int, default=30)
ap.add_argument("--tag", default="")
ap.add_argument("--regs", type=int, default=0)
a = ap.parse_args()
n = a.qubits
if a.regs:
    qi.engine.set_option("window_regs", a.regs)
st = qi.State.new_random(n)
full = 32.0 * (1 << n)
ctrl = n - 2
gates = {
    "h": (lambda t: st.h_(t), 1.0), "rx": (lambda t: st.rx_(t, 0.3), 1.0), "rz": (lambda t: st.rz_(t, 0.3), 1.0),
    "x": (lambda t: st.x_(t), 1.0), "p": (lambda t: st.p_(t, 0.3), 0.5),
    "cnot": (lambda t: st.cnot_(ctrl if t != ctrl else ctrl - 1, t), 0.5),
    "cp": (lambda t: st.cp_multi_([t], [ctrl if t != ctrl else ctrl - 1], 0.3), 0.25),
}
for name, (fn, frac) in gates.items():
    row = []
    for t in (0, 3, 5, 12, n - 1):
        fn(t)
        qi.engine.synchronize()
        qi.engine.timer_start()
        for _ in range(10):
            fn(t)
        ms = qi.engine.timer_stop() / 10
        row.append(f"t{t}:{full * frac / ms / 1e6:7.0f}GB/s({ms:.2f}ms)")
    print(a.tag, f"{name:5s}", " ".join(row), flush=True)
