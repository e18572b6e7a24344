"""Marginal cost of one fused op inside a window pass: passes with m identical gates (all in ONE launch)
at N qubits; the slope of time vs m is the per-op cost, the intercept the streaming cost."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quant_iron_b200 as qi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--qubits", type=int, default=30)
ap.add_argument("--tma", type=int, default=0)
ap.add_argument("--ms", default="1,8,16,32")
a = ap.parse_args()
n = a.qubits
qi.engine.set_option("tma", a.tma)
st = qi.State.new_random(n)
REG = [10, 14, 19, 25]
LANE = [0, 1, 2, 3]


def build(kind, m):
    b = qi.CircuitBuilder(n)
    for i in range(m):
        q = (LANE if kind.endswith("lane") else REG)[i % 4]
        if kind.startswith("h"):
            b.h_gate(q)
        elif kind.startswith("rx"):
            b.rx_gate(q, 0.1 + 0.01 * i)
        elif kind.startswith("ry"):
            b.ry_gate(q, 0.1 + 0.01 * i)
        elif kind.startswith("y"):
            b.y_gate(q)
        elif kind.startswith("cx"):
            b.cnot_gate(q, 7 + (i % 3))          # control on a tile qubit
        elif kind.startswith("ccx"):
            b.toffoli_gate(7, REG[(i + 1) % 4] if kind.endswith("reg") else 4, q)
    return b.build()


for kind in ("h_reg", "rx_reg", "y_reg", "cx_reg", "ccx_reg", "h_lane", "rx_lane", "cx_lane"):
    row = []
    for m in [int(x) for x in a.ms.split(",")]:
        c = build(kind, m)
        c.execute_(st)
        qi.engine.synchronize()
        qi.engine.stats_reset()
        qi.engine.timer_start()
        for _ in range(3):
            c.execute_(st)
        ms = qi.engine.timer_stop() / 3
        launches = sum(v["launches"] for v in qi.engine.stats().values()) // 3
        row.append(f"m={m}: {ms:.2f} ms ({launches} launch)")
    print(f"tma={a.tma} {kind:8s} " + "  ".join(row), flush=True)
