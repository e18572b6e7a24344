"""30-qubit single-gate passes per executor: interpreting tile kernel, JIT module (optionally with padded shared memory = capped
occupancy), warp-window kernel.  GB/s = algorithmic bytes / CUDA-event time over 20 back-to-back passes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quant_iron_b200 as qi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
qi.engine.init(0)
st = qi.State.new_random(n)
full = 32.0 * (1 << n)
u2 = qi.Unitary2.new([[0.6 + 0j, 0.8j], [0.8j, 0.6 + 0j]])
gates = {"h": lambda t: st.h_(t), "rx": lambda t: st.rx_(t, 0.3), "rz": lambda t: st.rz_(t, 0.3), "x": lambda t: st.x_(t),
         "u2": lambda t: st.apply_(u2, [t], [])}
configs = [("tile_interp", {"tile": 1, "jit": 0}), ("module", {"tile": 1, "jit": 2}), ("module_smem56", {"tile": 1, "jit": 2, "jit_smem_kb": 56}),
           ("module_smem76", {"tile": 1, "jit": 2, "jit_smem_kb": 76}), ("window", {"tile": 0, "jit": 0})]
for cname, opts in configs:
    for k, v in {"tile": 1, "jit": 1, "jit_smem_kb": 0}.items():
        qi.engine.set_option(k, v)
    for k, v in opts.items():
        qi.engine.set_option(k, v)
    for name, fn in gates.items():
        row = []
        for t in (0, 1, 4, 5, 12, n - 1):
            fn(t)
            fn(t)
            qi.engine.jit_drain()
            qi.engine.synchronize()
            qi.engine.stats_reset()
            qi.engine.timer_start()
            for _ in range(20):
                fn(t)
            ms = qi.engine.timer_stop() / 20
            row.append(f"t{t}:{full / ms / 1e6:6.0f}")
        print(f"{cname:14s} {name:3s}", " ".join(row), {k: v["launches"] for k, v in qi.engine.stats().items()}, flush=True)
