"""QFT on new_plus(n): time per executor (JIT tile modules, interpreting tile kernel, warp-window kernel).  python tools/prof_qft.py 30 [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quant_iron_b200 as qi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
qi.engine.init(0)
qft = qi.CircuitBuilder(n).add_subroutine(qi.Subroutine.qft(list(range(n)), n)).build()
for name, tile, jit in (("tile_jit", 1, 2), ("tile_interpreter", 1, 0), ("window", 0, 0)):
    qi.engine.set_option("tile", tile)
    qi.engine.set_option("jit", jit)
    st = qi.State.new_plus(n)
    qft.execute_(st)
    qi.engine.synchronize()
    a0 = st.amplitude(0)
    for _ in range(3):            # the QFT's trailing swaps are a relabelling: the layout (and the modules) alternate with period 2
        qft.execute_(st)
    qi.engine.synchronize()
    del st
    st = qi.State.new_plus(n)
    qi.engine.stats_reset()
    qi.engine.timer_start()
    for _ in range(reps):
        qft.execute_(st)
    ms = qi.engine.timer_stop() / reps
    print(f"n={n} {name}: {ms:.2f} ms per QFT, |amp0-1|={abs(a0 - 1):.2e}, kernels {({k: v['launches'] for k, v in qi.engine.stats().items()})} jit {qi.engine.jit_stats()}", flush=True)
    del st
