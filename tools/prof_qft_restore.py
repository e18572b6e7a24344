"""QFT on new_plus(n), modules only: per-execution times over the period-2 layout cycle, with and without tile_restore.  python tools/prof_qft_restore.py 33"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quant_iron_b200 as qi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 33
qi.engine.init(0)
qft = qi.CircuitBuilder(n).add_subroutine(qi.Subroutine.qft(list(range(n)), n)).build()
qi.engine.set_option("jit", 2)
for restore in (0, 1):
    qi.engine.set_option("tile_restore", restore)
    st = qi.State.new_plus(n)
    times = []
    for k in range(6):
        qi.engine.stats_reset()
        qi.engine.synchronize()
        qi.engine.timer_start()
        qft.execute_(st)
        ms = qi.engine.timer_stop()
        times.append((round(ms, 1), sum(v["launches"] for v in qi.engine.stats().values())))
    a0 = st.amplitude(0)
    print(f"n={n} tile_restore={restore}: (ms, launches) per execution {times}  final amp0={a0:.3e} norm={st.norm_sqr():.12f}", flush=True)
    del st
