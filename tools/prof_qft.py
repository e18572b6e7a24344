"""QFT on new_plus(n): time per option set (tile on/off), for profiling the phase-table ops.  python tools/prof_qft.py 30 [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quant_iron_b200 as qi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
qi.engine.init(0)
qft = qi.CircuitBuilder(n).add_subroutine(qi.Subroutine.qft(list(range(n)), n)).build()
for tile in (1, 0):
    qi.engine.set_option("tile", tile)
    st = qi.State.new_plus(n)
    qft.execute_(st)
    qi.engine.synchronize()
    st = qi.State.new_plus(n)
    qi.engine.stats_reset()
    qi.engine.timer_start()
    for _ in range(reps):
        qft.execute_(st)
    ms = qi.engine.timer_stop() / reps
    print(f"n={n} tile={tile}: {ms:.2f} ms per QFT, kernels {({k: v['launches'] for k, v in qi.engine.stats().items()})}", flush=True)
    del st
