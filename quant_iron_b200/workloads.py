"""Synthetic workloads shared by the GPU engine, the CPU oracle and bench.py.

Pure Python, no device access: a workload is a list of gate specs `(name, targets, controls, params)`
that `build_circuit` lowers through any implementation of the reference-shaped API (this package or
the oracle in tests/bench), so both sides execute the *same* circuit.
"""
from __future__ import annotations

import math
from typing import List, Tuple

MASK64 = (1 << 64) - 1
GOLDEN = 0x9E3779B97F4A7C15

SEED_CIRCUIT = 20260001
SEED_STATE = 20260002
SEED_MEASURE = 20260003

Spec = Tuple[str, list, list, list]


class SplitMix64:
    def __init__(self, seed: int):
        self.s = seed & MASK64

    def next_u64(self) -> int:
        self.s = (self.s + GOLDEN) & MASK64
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
        return z ^ (z >> 31)

    def uniform(self) -> float:
        return (self.next_u64() >> 11) * (1.0 / 9007199254740992.0)


def random_layered_circuit(num_qubits: int, depth: int = 40, seed: int = SEED_CIRCUIT) -> List[Spec]:
    """BASELINE.json config 2: each layer = one gate from {H, RX(t), RZ(t)} on every qubit
    (t = 2*pi*u), then brick-work CNOT(control q, target q+1) on pairs starting at `layer & 1`."""
    rng = SplitMix64(seed)
    specs: List[Spec] = []
    for layer in range(depth):
        for q in range(num_qubits):
            r = rng.next_u64() % 3
            if r == 0:
                specs.append(("h", [q], [], []))
            else:
                theta = 2.0 * math.pi * rng.uniform()
                specs.append(("rx" if r == 1 else "rz", [q], [], [theta]))
        for q in range(layer & 1, num_qubits - 1, 2):
            specs.append(("cnot", [q + 1], [q], []))      # target q+1, control q
    return specs


def qft_specs(num_qubits: int) -> List[Spec]:
    """Subroutine::qft over all qubits (subroutine.rs:90-112) as specs."""
    specs: List[Spec] = []
    for i in range(num_qubits):
        specs.append(("h", [i], [], []))
        den = 2.0
        for k in range(1, num_qubits - i):
            specs.append(("cp", [i], [i + k], [math.pi / den]))
            den *= 2.0
    for i in range(num_qubits // 2):
        specs.append(("swap", [i, num_qubits - 1 - i], [], []))
    return specs


def build_circuit(qi, num_qubits: int, specs: List[Spec]):
    """Lower specs through `qi.CircuitBuilder` (any implementation of the reference-shaped API)."""
    b = qi.CircuitBuilder(num_qubits)
    for name, targets, controls, params in specs:
        if name == "h":
            b.h_gate(targets[0])
        elif name == "x":
            b.x_gate(targets[0])
        elif name == "rx":
            b.rx_gate(targets[0], params[0])
        elif name == "ry":
            b.ry_gate(targets[0], params[0])
        elif name == "rz":
            b.rz_gate(targets[0], params[0])
        elif name == "p":
            b.p_gate(targets[0], params[0])
        elif name == "cp":
            b.cp_gates([targets[0]], controls, params[0])
        elif name == "cnot":
            b.cnot_gate(targets[0], controls[0])
        elif name == "swap":
            b.swap_gate(targets[0], targets[1])
        else:
            raise ValueError(f"unknown gate spec {name}")
    return b.build()


def algorithmic_bytes(num_qubits: int, specs: List[Spec]) -> float:
    """Sum over gates of the bytes an UNFUSED pass must move: 2*16*2^n*f."""
    full = 2.0 * 16.0 * float(1 << num_qubits)
    frac = {"h": 1.0, "x": 1.0, "rx": 1.0, "ry": 1.0, "rz": 1.0, "p": 0.5, "cp": 0.25, "cnot": 0.5, "swap": 0.5}
    return sum(full * frac[s[0]] for s in specs)
