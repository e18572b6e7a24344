"""Host-overhead diagnosis for small circuits (20-qubit QFT): where does the wall time go?"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quant_iron_b200 as qi  # noqa: E402

n = 20
qft = qi.CircuitBuilder(n).add_subroutine(qi.Subroutine.qft(list(range(n)), n)).build()
st = qi.State.new_plus(n)
qft.execute(st)
qi.engine.synchronize()


def wall(fn, reps=20):
    qi.engine.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    qi.engine.synchronize()
    return (time.perf_counter() - t0) * 1e3 / reps


print("clone only        ms", wall(lambda: st.clone()))
work = st.clone()
print("execute_ in place ms", wall(lambda: qft.execute_(work)))
print("execute (clone+)  ms", wall(lambda: qft.execute(st)))
for opt, val in (("absorb", 0), ("lazy_swap", 0), ("fuse", 0)):
    qi.engine.set_option(opt, val)
    print(f"execute_ with {opt}={val} ms", wall(lambda: qft.execute_(work)))
    qi.engine.set_option(opt, 1)
qi.engine.stats_reset()
qft.execute_(work)
print({k: v["launches"] for k, v in qi.engine.stats().items()})
h = qi.State.new_plus(n)
print("single h_ gate    ms", wall(lambda: h.h_(3)))
print("to_host           ms", wall(lambda: work.state_vector, reps=5))
