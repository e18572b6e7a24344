"""First-order Trotter steps of the Heisenberg chain (BASELINE config 3) + <H>: CUDA-event times.  python tools/prof_trotter.py [n] [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quant_iron_b200 as qi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
qi.engine.init(0)
h = qi.heisenberg_1d(n, 1.0, 2.0, 3.0, 0.5, 0.1)
st = qi.State.new_plus(n)
qi.trotter_evolve_state_(h, st, 0.01, 2, qi.TrotterOrder.First)
qi.engine.synchronize()
for rep in range(3):
    st = qi.State.new_plus(n)
    qi.engine.stats_reset()
    qi.engine.timer_start()
    qi.trotter_evolve_state_(h, st, 0.01, steps, qi.TrotterOrder.First)
    ms = qi.engine.timer_stop()
    qi.engine.timer_start()
    e = h.expectation_value(st)
    ms_e = qi.engine.timer_stop()
    print(f"n={n} steps={steps}: trotter {ms:.2f} ms, <H> {ms_e:.3f} ms, <H>={e.real:.12f}, norm={st.norm_sqr():.15f}, "
          f"kernels={ {k: v['launches'] for k, v in qi.engine.stats().items()} }", flush=True)
