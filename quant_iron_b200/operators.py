"""The `Operator` trait and the reference's operator types (src/components/operator.rs).

Each operator is a thin record: `apply(state, targets, controls)` clones the (borrowed) input state
on the device and runs the operator in place through `qi_apply_gate`; validation and its error
variants come from the C library (same order as validate_qubits, operator.rs:214-273).
User-defined operators subclass `Operator` and compose the primitives (see `Operator.record`).
"""
from __future__ import annotations

import math
from typing import List, Sequence

from . import _ffi
from .errors import Error

# qi_gate_kind (include/qiron_b200.h)
(G_H, G_X, G_Y, G_Z, G_I, G_S, G_SDG, G_T, G_TDG, G_P, G_RX, G_RY, G_RZ, G_U2, G_CNOT, G_SWAP, G_TOFFOLI,
 G_MATCHGATE) = range(1, 19)


def make_record(kind: int, targets: Sequence[int], controls: Sequence[int], params: Sequence[float]):
    """Build one `qi_gate` record; returns (record, keepalive)."""
    g = _ffi.QiGate()
    g.kind = kind
    ts = list(targets)
    g.num_targets = len(ts)
    for i, t in enumerate(ts[:2]):
        g.targets[i] = _as_u32(t)
    cs = [_as_u32(c) for c in controls]
    arr = _ffi.u32_array(cs)
    g.num_controls = len(cs)
    g.controls = arr
    for i, p in enumerate(list(params)[:8]):
        g.params[i] = float(p)
    return g, arr


def _as_u32(q) -> int:
    q = int(q)
    if q < 0:
        raise Error("InvalidArgument")
    return min(q, 0xFFFFFFFF)


class Operator:
    """operator.rs:151-190."""
    KIND = 0
    BASE_QUBITS = 1

    def params(self) -> List[float]:
        return []

    def record(self, targets, controls):
        """The C-ABI gate record(s) this operator stands for."""
        return make_record(self.KIND, targets, controls, self.params())

    def apply(self, state, target_qubits: Sequence[int], control_qubits: Sequence[int] = ()):
        out = state.clone()
        out.apply_(self, target_qubits, control_qubits)
        return out

    def base_qubits(self) -> int:
        return self.BASE_QUBITS

    def __repr__(self):
        return type(self).__name__


class Hadamard(Operator):
    KIND = G_H


class _PauliOp(Operator):
    def __init__(self, kind, name):
        self.KIND, self.name = kind, name

    def __repr__(self):
        return f"Pauli.{self.name}"

    def to_pauli_string(self, target_qubit: int):  # operator.rs:632-636
        from .pauli import PauliString
        return PauliString.with_ops(complex(1.0, 0.0), {target_qubit: self})


class Pauli:
    """`enum Pauli { X, Y, Z }` (operator.rs:443-450)."""
    X = _PauliOp(G_X, "X")
    Y = _PauliOp(G_Y, "Y")
    Z = _PauliOp(G_Z, "Z")


class CNOT(Operator):
    KIND, BASE_QUBITS = G_CNOT, 2


class SWAP(Operator):
    KIND, BASE_QUBITS = G_SWAP, 2


class Toffoli(Operator):
    KIND, BASE_QUBITS = G_TOFFOLI, 3


class Identity(Operator):
    KIND = G_I


class PhaseS(Operator):
    KIND = G_S


class PhaseT(Operator):
    KIND = G_T


class PhaseSdag(Operator):
    KIND = G_SDG


class PhaseTdag(Operator):
    KIND = G_TDG


class _Angle(Operator):
    def __init__(self, angle: float):
        self.angle = float(angle)

    @classmethod
    def new(cls, angle: float):
        return cls(angle)

    def params(self):
        return [self.angle]

    def __repr__(self):
        return f"{type(self).__name__}({self.angle:.3f})"


class PhaseShift(_Angle):
    KIND = G_P


class RotateX(_Angle):
    KIND = G_RX


class RotateY(_Angle):
    KIND = G_RY


class RotateZ(_Angle):
    KIND = G_RZ


def _cexp_i(phi: float) -> complex:
    # Complex::new(0.0, phi).exp() = e^0 * (cos phi, sin phi)   (num-complex)
    return complex(math.cos(phi), math.sin(phi))


class Unitary2(Operator):
    """operator.rs:2058-2275."""
    KIND = G_U2

    def __init__(self, matrix):
        self.matrix = [[complex(matrix[0][0]), complex(matrix[0][1])],
                       [complex(matrix[1][0]), complex(matrix[1][1])]]

    @staticmethod
    def new(matrix) -> "Unitary2":
        u = Unitary2(matrix)
        _ffi.check(_ffi.lib.qi_unitary2_check(_ffi.dbl_array(u.params())))   # operator.rs:2092-2118
        return u

    @staticmethod
    def from_ry_phase(theta: float, phi: float) -> "Unitary2":        # operator.rs:2140-2156
        ch, sh = math.cos(theta / 2.0), math.sin(theta / 2.0)
        e = _cexp_i(phi)
        return Unitary2([[complex(ch, 0.0), complex(-e.real * sh, -e.imag * sh)],
                         [complex(sh, 0.0), complex(e.real * ch, e.imag * ch)]])

    @staticmethod
    def from_ry_phase_dagger(theta: float, phi: float) -> "Unitary2":  # operator.rs:2173-2192
        ch, sh = math.cos(theta / 2.0), math.sin(theta / 2.0)
        e = _cexp_i(-phi)
        return Unitary2([[complex(ch, 0.0), complex(sh, 0.0)],
                         [complex(-e.real * sh, -e.imag * sh), complex(e.real * ch, e.imag * ch)]])

    def params(self):
        out = []
        for row in self.matrix:
            for z in row:
                out += [z.real, z.imag]
        return out


class Matchgate(Operator):
    """operator.rs:852-1019."""
    KIND, BASE_QUBITS = G_MATCHGATE, 2

    def __init__(self, theta: float, phi1: float, phi2: float):
        self.theta, self.phi1, self.phi2 = float(theta), float(phi1), float(phi2)

    @classmethod
    def new(cls, theta, phi1, phi2):
        return cls(theta, phi1, phi2)

    def params(self):
        return [self.theta, self.phi1, self.phi2]
