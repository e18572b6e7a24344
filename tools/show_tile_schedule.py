"""Host-only: the CTA-tile (k_tile) launches of a workload -- tile qubits, rounds, ops per round (no GPU needed)."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import quant_iron_b200 as qi  # noqa: E402
from quant_iron_b200 import workloads as w  # noqa: E402
import window_interp as wi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--qubits", type=int, default=30)
ap.add_argument("--depth", type=int, default=40)
ap.add_argument("--what", default="layered")
ap.add_argument("--lean", type=int, default=0)
ap.add_argument("-v", action="store_true")
a = ap.parse_args()
qi.engine.set_option("lean", a.lean)
specs = w.random_layered_circuit(a.qubits, a.depth) if a.what == "layered" else w.qft_specs(a.qubits)
steps, arena, phys = wi.parse(wi.lower(w.build_circuit(qi, a.qubits, specs), a.qubits))
tiles = [s for s in steps if s[0] == "tile"]
tot_rounds = tot_ops = 0
print(f"{len(specs)} gates -> {len(steps)} steps ({len(tiles)} tile launches)")
for i, s in enumerate(steps):
    if s[0] != "tile":
        print(i, s[0])
        continue
    rounds = s[2]
    nops = [len(r[2]) for r in rounds]
    kinds = np.concatenate([r[2]["kind"] for r in rounds]) if sum(nops) else np.zeros(0, dtype=np.uint8)
    pair = int(np.isin(kinds, list(wi.PAIR_KINDS)).sum())
    x = int((kinds == wi.WK_X).sum())
    tab = int((kinds == wi.WK_TABLE).sum())
    tot_rounds += len(rounds)
    tot_ops += sum(nops)
    moved = sum(1 for a_, b_ in zip(s[1], s[3]) if a_ != b_)
    print(f"{i:3d} window={s[1][5:]} moved={moved} rounds={len(rounds):2d} ops={sum(nops):3d} (pair {pair}, x {x}, table {tab}, other {len(kinds) - pair - tab}) per-round={nops}")
    if a.v:
        for regs, thr, ops in rounds:
            print("      regs", regs, "thr", thr, "kinds", list(ops["kind"]))
print(f"total rounds {tot_rounds}, ops {tot_ops}")
