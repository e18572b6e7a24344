"""Kernel experiment driver: times the layered circuit and the QFT through whatever library
QIRON_B200_LIB points at (default: the in-tree build)."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quant_iron_b200 as qi  # noqa: E402
from quant_iron_b200 import workloads as w  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--qubits", type=int, default=30)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--tag", default="")
ap.add_argument("--regs", type=int, default=0)
ap.add_argument("--tma", type=int, default=0)
ap.add_argument("--absorb", type=int, default=1)
a = ap.parse_args()
n = a.qubits
if a.regs:
    qi.engine.set_option("window_regs", a.regs)
qi.engine.set_option("tma", a.tma)
qi.engine.set_option("absorb", a.absorb)
st = qi.State.new_zero(n)
for name, specs in (("layered", w.random_layered_circuit(n, 40)), ("qft", w.qft_specs(n))):
    c = w.build_circuit(qi, n, specs)
    c.execute_(st)
    qi.engine.synchronize()
    qi.engine.stats_reset()
    t0 = time.perf_counter()
    qi.engine.timer_start()
    for _ in range(a.reps):
        c.execute_(st)
    ms = qi.engine.timer_stop() / a.reps
    host_ms = (time.perf_counter() - t0) * 1e3 / a.reps
    stats = qi.engine.stats()
    print(f"{a.tag} n={n} {name}: {ms:.1f} ms/circuit ({len(specs) / ms * 1e3:.0f} gates/s), host {host_ms:.1f} ms, "
          f"launches/circuit={ {k: v['launches'] // a.reps for k, v in stats.items()} }", flush=True)
print(a.tag, "norm", st.norm_sqr())
