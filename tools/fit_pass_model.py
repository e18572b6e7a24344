"""Per-pass cost model of the fused window kernel, fitted on a committed launch list (no GPU needed).

    python tools/fit_pass_model.py [profiles/r01_launches_bench_final.csv]

The launch list (ncu --metrics gpu__time_duration.sum of `bench.py --steps 2 --warmup 1`) holds one duration per
k_window launch; the host-only scheduler (qi_debug_schedule) says what each of those passes contained when the list was
taken (option late_tables = 0: the list predates the late table placement).  Model per pass:
    comp = base + c_lane * lane_ops + c_reg * register_ops + c_diag * diagonal_ops + c_half * absorbed_halves
    t    = (floor^4 + comp^4)^(1/4)        (smooth maximum of the HBM floor and the op time)
Prints the fitted coefficients, the fit error, and what the model says about the current schedule variants."""
import csv
import os
import sys

import numpy as np
from scipy.optimize import least_squares

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import quant_iron_b200 as qi  # noqa: E402
from quant_iron_b200 import workloads as w  # noqa: E402
from show_schedule import schedule  # noqa: E402


def model(p, S):
    floor, base, c_lane, c_reg, c_diag, c_half = p
    comp = base + c_lane * S[:, 2] + c_reg * (S[:, 3] - S[:, 6]) + c_diag * S[:, 5] + c_half * S[:, 6]
    return (floor ** 4 + np.maximum(comp, 0.0) ** 4) ** 0.25


def main():
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r01_launches_bench_final.csv")
    ms = []
    with open(path) as f:
        for line in f:
            if line.startswith('"') and "k_window" in line:
                ms.append(float(next(csv.reader([line]))[-1]) / 1e6)
    n = 30
    specs = w.random_layered_circuit(n, 40)
    qi.engine.set_option("late_tables", 0)
    S = np.array(schedule(n, specs), float)
    reps = len(ms) // len(S)
    T = np.array(ms[:reps * len(S)]).reshape(reps, len(S))[1:].mean(axis=0)
    fit = least_squares(lambda p: model(p, S) - T, [5.2, 1.0, 0.75, 0.4, 0.7, 0.15])
    p = fit.x
    print(f"launch list: {path}: {len(ms)} k_window launches = {reps} circuits x {len(S)} passes, {T.sum():.1f} ms per circuit")
    print("fit (ms at 30 qubits): floor %.2f  base %.2f  lane op %.2f  register op %.2f  diagonal op %.2f  absorbed half %.2f" % tuple(p))
    print(f"mean |error| per pass {np.abs(model(p, S) - T).mean():.2f} ms; model total {model(p, S).sum():.1f} ms")
    comp = p[1] + p[2] * S[:, 2] + p[3] * (S[:, 3] - S[:, 6]) + p[4] * S[:, 5] + p[5] * S[:, 6]
    print(f"sum of op time {comp.sum():.0f} ms, sum of HBM floors {len(S) * p[0]:.0f} ms, perfect-overlap bound {np.maximum(comp, p[0]).sum():.0f} ms")
    for name, opts, regs in (("as measured (late_tables=0)", {"late_tables": 0}, 4), ("late_tables=1 (default now)", {"late_tables": 1}, 4),
                             ("late_tables=1, window_regs=5 (same coefficients assumed)", {"late_tables": 1}, 5)):
        for k, v in opts.items():
            qi.engine.set_option(k, v)
        R = np.array(schedule(n, specs, regs), float)
        print(f"  {name}: {len(R)} passes, {int(R[:, 5].sum())} diagonal ops, model {model(p, R).sum():.1f} ms")
    qi.engine.set_option("late_tables", 1)


if __name__ == "__main__":
    main()
