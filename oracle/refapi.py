"""CPU ORACLE -- reference-shaped API over oracle/qi_oracle.c.

TEST INFRASTRUCTURE ONLY (see the header of qi_oracle.c).  Only tests/,
__graft_entry__.smoke() and bench.py's CPU-baseline legs import this module.

It restates, on host numpy arrays, the slice of quant-iron's public surface that
sits on the state-vector hot path, with the reference's names, argument order
and error behaviour, so that the reference's own tests can be replayed against
it (tests/test_ref_ported_*.py) and so that the GPU engine can be compared with
it call for call.  Citations are file:line under the reference root.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Iterable, List, Optional, Sequence

import numpy as np

from . import build as _build

_lib = C.CDLL(_build.build())

_u32p = C.POINTER(C.c_uint32)
_i32p = C.POINTER(C.c_int32)
_dp = C.POINTER(C.c_double)
_vp = C.c_void_p

_lib.orc_gate.argtypes = [_vp, C.c_int, C.c_int, _u32p, C.c_int, _u32p, C.c_int, _dp]
_lib.orc_gate.restype = None
_lib.orc_gate_faithful.argtypes = [_vp, _vp, C.c_int, C.c_int, _u32p, C.c_int, _u32p, C.c_int, _dp]
_lib.orc_gate_faithful.restype = None
_lib.orc_scale.argtypes = [_vp, C.c_int64, _dp]
_lib.orc_add.argtypes = [_vp, _vp, C.c_int64]
_lib.orc_sub.argtypes = [_vp, _vp, C.c_int64]
_lib.orc_inner_product.argtypes = [_vp, _vp, C.c_int64, _dp]
_lib.orc_norm_sqr.argtypes = [_vp, C.c_int64]
_lib.orc_norm_sqr.restype = C.c_double
_lib.orc_normalise.argtypes = [_vp, C.c_int64]
_lib.orc_normalise.restype = C.c_int
_lib.orc_pauli_apply.argtypes = [_vp, C.c_int, _u32p, _i32p, C.c_int, _dp]
_lib.orc_pauli_exp.argtypes = [_vp, C.c_int, _u32p, _i32p, C.c_int, _dp]
_lib.orc_pauli_expect.argtypes = [_vp, C.c_int, _u32p, _i32p, C.c_int, _dp, _dp]
_lib.orc_probabilities.argtypes = [_vp, C.c_int, _u32p, C.c_int, _dp]
_lib.orc_sample_bin.argtypes = [_dp, C.c_int64, C.c_double]
_lib.orc_sample_bin.restype = C.c_int64
_lib.orc_sample_margin.argtypes = [_dp, C.c_int64, C.c_double]
_lib.orc_sample_margin.restype = C.c_double
_lib.orc_collapse.argtypes = [_vp, C.c_int, _u32p, C.c_int, C.c_uint64]
_lib.orc_uniform.argtypes = [C.c_uint64, C.c_uint64]
_lib.orc_uniform.restype = C.c_double
_lib.orc_sample.argtypes = [_dp, C.c_int64, C.c_uint64, C.c_int64, C.POINTER(C.c_int64)]
_lib.orc_random_state.argtypes = [_vp, C.c_int, C.c_uint64]
_lib.orc_num_threads.restype = C.c_int
for _f in ("orc_c_exp", "orc_c_cosh", "orc_c_sinh"):
    getattr(_lib, _f).argtypes = [_dp, _dp]

# gate kinds (same values as qi_oracle.c)
G_H, G_X, G_Y, G_Z, G_I, G_S, G_SDG, G_T, G_TDG, G_P, G_RX, G_RY, G_RZ, G_U2, \
    G_CNOT, G_SWAP, G_TOFFOLI, G_MATCHGATE = range(1, 19)

F32_EPS = 1.1920928955078125e-07   # f32::EPSILON, State::eq tolerance (state.rs:2360-2367)
F64_EPS = 2.220446049250313e-16

PARALLEL_THRESHOLD_NUM_QUBITS = 10  # operator.rs:18


def num_threads() -> int:
    return int(_lib.orc_num_threads())


def _u32(xs: Sequence[int]):
    arr = (C.c_uint32 * max(1, len(xs)))(*xs)
    return arr


def _dbl(xs: Sequence[float]):
    return (C.c_double * max(1, len(xs)))(*xs)


def _cfun(name: str, z: complex) -> complex:
    out = (C.c_double * 2)()
    getattr(_lib, name)(_dbl([z.real, z.imag]), out)
    return complex(out[0], out[1])


def c_exp(z: complex) -> complex:
    return _cfun("orc_c_exp", z)


def c_cosh(z: complex) -> complex:
    return _cfun("orc_c_cosh", z)


def c_sinh(z: complex) -> complex:
    return _cfun("orc_c_sinh", z)


def c_mul(a: complex, b: complex) -> complex:
    """num-complex product, plain f64 ops (Python's own complex `*` agrees for finite values)."""
    return complex(a.real * b.real - a.imag * b.imag, a.real * b.imag + a.imag * b.real)


def uniform(seed: int, k: int = 0) -> float:
    """k-th draw of the shared-seed stream (splitmix64, u = (x >> 11) * 2^-53)."""
    return float(_lib.orc_uniform(C.c_uint64(seed & (2**64 - 1)), C.c_uint64(k)))


# ----------------------------------------------------------------------------
class Error(Exception):
    """errors.rs:3-97.  `variant` is the Rust variant name, `payload` its fields."""

    def __init__(self, variant: str, *payload):
        super().__init__(f"{variant}{payload}")
        self.variant = variant
        self.payload = tuple(payload)

    # thiserror Display strings (errors.rs:10-98)
    _DISPLAY = {
        "InvalidNumberOfMeasurements": "Invalid number of measurements: {0}",
        "OverlappingControlAndTargetQubits": "Control qubit index {0} overlaps with target qubit index {1}",
        "InvalidNumberOfQubits": "Invalid number of qubits: {0}",
        "InvalidQubitIndex": "Invalid qubit index: {0} for {1} qubits",
        "StateVectorNotNormalised": "State vector is not normalised",
        "NonUnitaryMatrix": "Non-unitary matrix",
        "InvalidNumberOfInputs": "Unexpected number of inputs: expected {1}, got {0}",
        "MismatchedNumberOfParameters": "Mismatched number of parameters: expected {0}, got {1}",
        "UnknownError": "An unknown error occurred",
        "CircuitMacroError": "Failed to create circuit from macro: {0}",
        "InvalidInputValue": "Invalid input value for operation: {0}",
        "ZeroNorm": "The state cannot be normalised because it has zero norm.",
        "InvalidPauliStringCoefficient": "Invalid Pauli String coefficient: {0}",
    }

    def to_string(self) -> str:
        fmt = self._DISPLAY.get(self.variant)
        try:
            return fmt.format(*self.payload) if fmt else str(self)
        except IndexError:
            return str(self)

    def __eq__(self, other):
        return isinstance(other, Error) and (self.variant, self.payload) == (other.variant, other.payload)

    def __hash__(self):
        return hash((self.variant, self.payload))


def validate_qubits(state: "State", targets: Sequence[int], controls: Sequence[int], expected: int):
    """operator.rs:214-273, same order of checks so the same variant fires."""
    if len(targets) != expected:
        raise Error("InvalidNumberOfQubits", len(targets))
    n = state.num_qubits
    for t in targets:
        if t >= n:
            raise Error("InvalidQubitIndex", t, n)
    for c in controls:
        if c >= n:
            raise Error("InvalidQubitIndex", c, n)
        for t in targets:
            if c == t:
                raise Error("OverlappingControlAndTargetQubits", c, t)
    if expected > 1:
        seen = set()
        for t in targets:
            if t in seen:
                raise Error("InvalidQubitIndex", t, n)
            seen.add(t)


# ----------------------------------------------------------------------------
class Operator:
    """operator.rs:151-190."""
    KIND = 0
    EXPECTED_TARGETS = 1
    BASE_QUBITS = 1

    def params(self) -> List[float]:
        return []

    def extra_validate(self, state, targets, controls):
        pass

    def apply(self, state: "State", target_qubits: Sequence[int], control_qubits: Sequence[int] = ()) -> "State":
        targets, controls = list(target_qubits), list(control_qubits)
        validate_qubits(state, targets, controls, self.EXPECTED_TARGETS)
        self.extra_validate(state, targets, controls)
        out = state.clone()
        _lib.orc_gate(out._ptr(), state.num_qubits, self.KIND, _u32(targets), len(targets),
                      _u32(controls), len(controls), _dbl(self.params()))
        return out

    def base_qubits(self) -> int:
        return self.BASE_QUBITS


class Hadamard(Operator):
    KIND = G_H


class _PauliOp(Operator):
    def __init__(self, kind, name):
        self.KIND, self.name = kind, name

    def __repr__(self):
        return f"Pauli.{self.name}"

    def to_pauli_string(self, target_qubit: int) -> "PauliString":  # operator.rs:632-636
        return PauliString.with_ops(complex(1.0, 0.0), {target_qubit: self})


class Pauli:
    X = _PauliOp(G_X, "X")
    Y = _PauliOp(G_Y, "Y")
    Z = _PauliOp(G_Z, "Z")


class CNOT(Operator):  # operator.rs:645-685
    KIND, BASE_QUBITS = G_CNOT, 2

    def extra_validate(self, state, targets, controls):
        if len(controls) != 1:
            raise Error("InvalidNumberOfQubits", len(controls))


class SWAP(Operator):  # operator.rs:731-820
    KIND, EXPECTED_TARGETS, BASE_QUBITS = G_SWAP, 2, 2


class Toffoli(Operator):  # operator.rs:1055-1075
    KIND, BASE_QUBITS = G_TOFFOLI, 3

    def extra_validate(self, state, targets, controls):
        if len(controls) != 2:
            raise Error("InvalidNumberOfQubits", len(controls))
        if controls[0] == controls[1]:
            raise Error("InvalidNumberOfQubits", len(controls))


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


class PhaseShift(_Angle):
    KIND = G_P


class RotateX(_Angle):
    KIND = G_RX


class RotateY(_Angle):
    KIND = G_RY


class RotateZ(_Angle):
    KIND = G_RZ


class Unitary2(Operator):  # operator.rs:2058-2275
    KIND = G_U2

    def __init__(self, matrix, _unchecked=False):
        self.matrix = [[complex(matrix[0][0]), complex(matrix[0][1])],
                       [complex(matrix[1][0]), complex(matrix[1][1])]]

    @staticmethod
    def new(matrix) -> "Unitary2":  # operator.rs:2092-2118
        u = Unitary2(matrix)
        tol = F64_EPS * 2.0
        (a, b), (c, d) = u.matrix

        def nsq(z):
            return z.real * z.real + z.imag * z.imag
        if abs((nsq(a) + nsq(b)) - 1.0) > tol:
            raise Error("NonUnitaryMatrix")
        if abs((nsq(c) + nsq(d)) - 1.0) > tol:
            raise Error("NonUnitaryMatrix")
        dot = c_mul(a, c.conjugate()) + c_mul(b, d.conjugate())
        if nsq(dot) > tol * tol:
            raise Error("NonUnitaryMatrix")
        return u

    @staticmethod
    def from_ry_phase(theta: float, phi: float) -> "Unitary2":  # operator.rs:2140-2156
        ch, sh = math.cos(theta / 2.0), math.sin(theta / 2.0)
        e = c_exp(complex(0.0, phi))
        return Unitary2([[complex(ch, 0.0), complex(-e.real * sh, -e.imag * sh)],
                         [complex(sh, 0.0), complex(e.real * ch, e.imag * ch)]])

    @staticmethod
    def from_ry_phase_dagger(theta: float, phi: float) -> "Unitary2":  # operator.rs:2173-2192
        ch, sh = math.cos(theta / 2.0), math.sin(theta / 2.0)
        e = c_exp(complex(0.0, -phi))
        return Unitary2([[complex(ch, 0.0), complex(sh, 0.0)],
                         [complex(-e.real * sh, -e.imag * sh), complex(e.real * ch, e.imag * ch)]])

    def params(self):
        out = []
        for row in self.matrix:
            for z in row:
                out += [z.real, z.imag]
        return out


class Matchgate(Operator):  # operator.rs:893-1014
    KIND, BASE_QUBITS = G_MATCHGATE, 2

    def __init__(self, theta: float, phi1: float, phi2: float):
        self.theta, self.phi1, self.phi2 = float(theta), float(phi1), float(phi2)

    @classmethod
    def new(cls, theta, phi1, phi2):
        return cls(theta, phi1, phi2)

    def extra_validate(self, state, targets, controls):
        if targets[0] == state.num_qubits - 1:  # operator.rs:903-905
            raise Error("InvalidQubitIndex", targets[0], state.num_qubits)

    def params(self):
        return [self.theta, self.phi1, self.phi2]


# ----------------------------------------------------------------------------
class MeasurementBasis:
    """measurement.rs:76-86."""

    def __init__(self, name: str, matrix=None):
        self.name, self.matrix = name, matrix

    @staticmethod
    def Custom(matrix) -> "MeasurementBasis":
        return MeasurementBasis("Custom", [[complex(z) for z in row] for row in matrix])

    def __eq__(self, other):
        return isinstance(other, MeasurementBasis) and self.name == other.name and self.matrix == other.matrix

    def __repr__(self):
        return f"MeasurementBasis.{self.name}"


MeasurementBasis.Computational = MeasurementBasis("Computational")
MeasurementBasis.X = MeasurementBasis("X")
MeasurementBasis.Y = MeasurementBasis("Y")


class MeasurementResult:
    """measurement.rs:15-25; attribute access falls through to new_state (Deref, 28-34)."""

    def __init__(self, basis, indices, outcomes, new_state):
        self.basis, self.indices, self.outcomes, self.new_state = basis, list(indices), list(outcomes), new_state

    def get_indices(self):
        return self.indices

    def get_basis(self):
        return self.basis

    def get_outcomes(self):
        return self.outcomes

    def get_new_state(self):
        return self.new_state

    def __getattr__(self, name):
        return getattr(self.new_state, name)


def _adjoint(m):  # state.rs:17-30
    return [[m[0][0].conjugate(), m[1][0].conjugate()], [m[0][1].conjugate(), m[1][1].conjugate()]]


# ----------------------------------------------------------------------------
class State:
    """state.rs:74-81 -- host numpy complex128 vector + qubit count."""

    def __init__(self, state_vector, num_qubits: int):
        # equivalent of the struct literal State { state_vector, num_qubits } (no checks)
        self.state_vector = np.ascontiguousarray(state_vector, dtype=np.complex128)
        self.num_qubits = int(num_qubits)

    def _ptr(self):
        return self.state_vector.ctypes.data_as(C.c_void_p)

    def clone(self) -> "State":
        return State(self.state_vector.copy(), self.num_qubits)

    # ---- constructors ----
    @staticmethod
    def new(state_vector) -> "State":  # state.rs:99-127
        v = np.ascontiguousarray(state_vector, dtype=np.complex128)
        ln = v.shape[0]
        if ln == 0:
            raise Error("InvalidNumberOfQubits", 0)
        if ln & (ln - 1):
            raise Error("InvalidNumberOfQubits", int(math.floor(math.log2(ln))))
        n = ln.bit_length() - 1
        norm = float(_lib.orc_norm_sqr(v.ctypes.data_as(C.c_void_p), ln))
        if abs(norm - 1.0) > F64_EPS * ln:
            raise Error("StateVectorNotNormalised")
        return State(v, n)

    @staticmethod
    def new_hartree_fock(num_electrons: int, num_orbitals: int) -> "State":  # state.rs:140-151
        if num_orbitals == 0 or num_orbitals < num_electrons:
            raise Error("InvalidInputValue", num_orbitals)
        n = ((1 << num_electrons) - 1) << (num_orbitals - num_electrons)
        return State.new_basis_n(num_orbitals, n)

    @staticmethod
    def new_zero(num_qubits: int) -> "State":  # state.rs:165-178
        if num_qubits == 0:
            raise Error("InvalidNumberOfQubits", 0)
        v = np.zeros(1 << num_qubits, dtype=np.complex128)
        v[0] = 1.0
        return State(v, num_qubits)

    @staticmethod
    def new_basis_n(num_qubits: int, n: int) -> "State":  # state.rs:194-210
        dim = 1 << num_qubits
        if n >= dim:
            raise Error("InvalidQubitIndex", n, num_qubits)
        if num_qubits == 0:
            raise Error("InvalidNumberOfQubits", 0)
        v = np.zeros(dim, dtype=np.complex128)
        v[n] = 1.0
        return State(v, num_qubits)

    @staticmethod
    def new_plus(num_qubits: int) -> "State":  # state.rs:225-237
        if num_qubits == 0:
            raise Error("InvalidNumberOfQubits", 0)
        dim = 1 << num_qubits
        return State(np.full(dim, 1.0 / math.sqrt(float(dim)), dtype=np.complex128), num_qubits)

    @staticmethod
    def new_minus(num_qubits: int) -> "State":  # state.rs:252-290
        if num_qubits == 0:
            raise Error("InvalidNumberOfQubits", 0)
        dim = 1 << num_qubits
        amp = 1.0 / math.sqrt(float(dim))
        idx = np.arange(dim, dtype=np.uint64)
        par = np.zeros(dim, dtype=np.uint64)
        for b in range(num_qubits):
            par ^= (idx >> np.uint64(b)) & np.uint64(1)
        v = np.where(par == 0, amp, -amp).astype(np.complex128)
        return State(v, num_qubits)

    @staticmethod
    def new_ghz(num_qubits: int) -> "State":  # state.rs:305-325
        if num_qubits == 0:
            raise Error("InvalidNumberOfQubits", 0)
        dim = 1 << num_qubits
        v = np.zeros(dim, dtype=np.complex128)
        v[0] = v[dim - 1] = math.sqrt(0.5)  # FRAC_1_SQRT_2
        return State(v, num_qubits)

    @staticmethod
    def _bell(v):
        return State(np.array(v, dtype=np.complex128), 2)

    @staticmethod
    def new_phi_plus():  # state.rs:31-66 bell_vectors
        a = math.sqrt(0.5)  # FRAC_1_SQRT_2
        return State._bell([a, 0, 0, a])

    @staticmethod
    def new_phi_minus():
        a = math.sqrt(0.5)  # FRAC_1_SQRT_2
        return State._bell([a, 0, 0, -a])

    @staticmethod
    def new_psi_plus():
        a = math.sqrt(0.5)  # FRAC_1_SQRT_2
        return State._bell([0, a, a, 0])

    @staticmethod
    def new_psi_minus():
        a = math.sqrt(0.5)  # FRAC_1_SQRT_2
        return State._bell([0, a, -a, 0])

    # ---- accessors ----
    def equals_without_phase(self, other: "State") -> bool:  # state.rs:384-390
        if self.num_qubits != other.num_qubits:
            return False
        return abs(abs(self.inner_product(other)) - 1.0) < F32_EPS

    def conj(self) -> "State":
        return State(np.conj(self.state_vector), self.num_qubits)

    def probability(self, n: int) -> float:  # state.rs:418-424
        if n >= self.state_vector.shape[0]:
            raise Error("InvalidQubitIndex", n, self.num_qubits)
        z = self.state_vector[n]
        return z.real * z.real + z.imag * z.imag

    def amplitude(self, n: int) -> complex:  # state.rs:448-453
        if n >= self.state_vector.shape[0]:
            raise Error("InvalidQubitIndex", n, self.num_qubits)
        return complex(self.state_vector[n])

    def fs_dist(self, other):  # state.rs:470-480
        return math.acos(abs(self.normalise().inner_product(other.normalise())))

    def fs_fidelity(self, other):  # state.rs:492-498
        return abs(self.normalise().inner_product(other.normalise())) ** 2

    def __eq__(self, other):  # state.rs:2348-2372
        if not isinstance(other, State) or self.num_qubits != other.num_qubits:
            return False
        if self.state_vector.shape != other.state_vector.shape:
            return False
        d = self.state_vector - other.state_vector
        return bool(np.all(np.abs(d.real) <= F32_EPS) and np.all(np.abs(d.imag) <= F32_EPS))

    __hash__ = None

    # ---- linear algebra ----
    def inner_product(self, other: "State") -> complex:  # state.rs:890-917
        if self.num_qubits == 0 or other.num_qubits == 0:
            raise Error("InvalidNumberOfQubits", 0)
        if self.state_vector.shape[0] != other.state_vector.shape[0]:
            raise Error("InvalidNumberOfQubits", self.num_qubits)
        out = (C.c_double * 2)()
        _lib.orc_inner_product(self._ptr(), other._ptr(), self.state_vector.shape[0], out)
        return complex(out[0], out[1])

    def normalise(self) -> "State":  # state.rs:924-945
        out = self.clone()
        if _lib.orc_normalise(out._ptr(), out.state_vector.shape[0]) != 0:
            raise Error("ZeroNorm")
        return out

    def tensor_product(self, other: "State") -> "State":  # state.rs:801-836 (self in the HIGH bits)
        if self.num_qubits == 0 or other.num_qubits == 0:
            raise Error("InvalidNumberOfQubits", 0)
        return State.new(np.kron(self.state_vector, other.state_vector))

    def __mul__(self, rhs):  # state.rs:2687-2726  amplitude * rhs
        z = complex(rhs)
        out = self.clone()
        _lib.orc_scale(out._ptr(), out.state_vector.shape[0], _dbl([z.real, z.imag]))
        return out

    __rmul__ = __mul__   # state.rs:2728-2776 (complex multiplication commutes bit-for-bit)

    def __add__(self, rhs: "State"):  # state.rs:2779-2800
        if self.num_qubits != rhs.num_qubits:
            raise RuntimeError("Cannot add states with different numbers of qubits")
        out = self.clone()
        _lib.orc_add(out._ptr(), rhs._ptr(), out.state_vector.shape[0])
        return out

    def __sub__(self, rhs: "State"):  # state.rs:2841-2862
        if self.num_qubits != rhs.num_qubits:
            raise RuntimeError("Cannot subtract states with different numbers of qubits")
        out = self.clone()
        _lib.orc_sub(out._ptr(), rhs._ptr(), out.state_vector.shape[0])
        return out

    # ---- operate (state.rs:970-1002) ----
    def operate(self, unitary: Operator, target_qubits, control_qubits=()):
        nt, nc = len(target_qubits), len(control_qubits)
        if unitary.base_qubits() != nt + nc:
            raise Error("InvalidNumberOfQubits", unitary.base_qubits())
        if nt > self.num_qubits:
            raise Error("InvalidNumberOfQubits", self.num_qubits)
        for q in list(target_qubits) + list(control_qubits):
            if q >= self.num_qubits:
                raise Error("InvalidQubitIndex", q, self.num_qubits)
        return unitary.apply(self, target_qubits, control_qubits)

    # ---- gate helpers ----
    def _multi(self, op: Operator, targets, controls=()):
        s = self.clone()
        for q in targets:
            s = op.apply(s, [q], controls)
        return s

    # ---- measurement (state.rs:525-784) ----
    def _check_measured(self, measured_qubits):
        actual = list(range(self.num_qubits)) if len(measured_qubits) == 0 else list(measured_qubits)
        if len(actual) > self.num_qubits:
            raise Error("InvalidNumberOfQubits", self.num_qubits)
        for q in actual:
            if q >= self.num_qubits:
                raise Error("InvalidQubitIndex", q, self.num_qubits)
        return actual

    def probabilities(self, qubits) -> np.ndarray:
        """un-normalised marginal table, bin bit j <-> qubits[j] (state.rs:559-588)."""
        probs = np.zeros(1 << len(qubits), dtype=np.float64)
        _lib.orc_probabilities(self._ptr(), self.num_qubits, _u32(qubits), len(qubits),
                               probs.ctypes.data_as(_dp))
        return probs

    def _measure_u(self, basis, qubits, u: float) -> MeasurementResult:
        if basis.name == "Computational":
            probs = self.probabilities(qubits)
            b = int(_lib.orc_sample_bin(probs.ctypes.data_as(_dp), probs.shape[0], u))
            if b < 0:
                raise Error("UnknownError")
            new = self.clone()
            _lib.orc_collapse(new._ptr(), self.num_qubits, _u32(qubits), len(qubits), b)
            outcomes = [(b >> j) & 1 for j in range(len(qubits))]
            return MeasurementResult(basis, qubits, outcomes, State.new(new.state_vector))
        if basis.name == "X":  # state.rs:670-686
            r = self.h_multi(qubits)._measure_u(MeasurementBasis.Computational, qubits, u)
            return MeasurementResult(basis, r.indices, r.outcomes, r.new_state.h_multi(qubits))
        if basis.name == "Y":  # state.rs:687-705
            r = self.s_dag_multi(qubits).h_multi(qubits)._measure_u(MeasurementBasis.Computational, qubits, u)
            return MeasurementResult(basis, r.indices, r.outcomes, r.new_state.h_multi(qubits).s_multi(qubits))
        # Custom: U before, U^dagger after (state.rs:706-728)
        r = self.unitary_multi(qubits, basis.matrix)._measure_u(MeasurementBasis.Computational, qubits, u)
        return MeasurementResult(basis, r.indices, r.outcomes,
                                 r.new_state.unitary_multi(qubits, _adjoint(basis.matrix)))

    def measure(self, basis, measured_qubits=(), seed: Optional[int] = None) -> MeasurementResult:
        """state.rs:525-730.  The reference draws from an unseedable thread RNG; the shared-seed
        contract of this build makes the draw u = uniform(seed, 0)."""
        qubits = self._check_measured(measured_qubits)
        if seed is None:
            seed = int.from_bytes(os.urandom(8), "little")
        return self._measure_u(basis, qubits, uniform(seed, 0))

    def measure_n(self, basis, measured_qubits, n: int, seed: Optional[int] = None) -> List[MeasurementResult]:
        """state.rs:750-784; shot k uses u_k = uniform(seed, k)."""
        if n == 0:
            raise Error("InvalidNumberOfMeasurements", 0)
        qubits = self._check_measured(measured_qubits)
        if seed is None:
            seed = int.from_bytes(os.urandom(8), "little")
        return [self._measure_u(basis, qubits, uniform(seed, k)) for k in range(n)]

    def sample_counts(self, measured_qubits, shots: int, seed: int):
        """bins of `shots` computational-basis draws from one probability table (what measure_n's
        outcomes are, without materialising the collapsed states)."""
        qubits = self._check_measured(measured_qubits)
        probs = self.probabilities(qubits)
        bins = np.zeros(shots, dtype=np.int64)
        _lib.orc_sample(probs.ctypes.data_as(_dp), probs.shape[0], C.c_uint64(seed), shots,
                        bins.ctypes.data_as(C.POINTER(C.c_int64)))
        return bins

    def sample_margin(self, measured_qubits, shots: int, seed: int) -> float:
        qubits = self._check_measured(measured_qubits)
        probs = self.probabilities(qubits)
        return min(float(_lib.orc_sample_margin(probs.ctypes.data_as(_dp), probs.shape[0], uniform(seed, k)))
                   for k in range(shots))


def _install_gate_methods():
    """state.rs:1019-2345: the ~70 convenience methods, generated from a table."""
    simple = {  # name -> operator factory (no parameters)
        "h": Hadamard, "x": lambda: Pauli.X, "y": lambda: Pauli.Y, "z": lambda: Pauli.Z, "i": Identity,
        "s": PhaseS, "t": PhaseT, "s_dag": PhaseSdag, "t_dag": PhaseTdag,
    }
    for name, fac in simple.items():
        def single(self, index, _f=fac):
            return _f().apply(self, [index], [])

        def multi(self, qubits, _f=fac):
            return self._multi(_f(), qubits)

        def cmulti(self, target_qubits, control_qubits, _f=fac):
            return self._multi(_f(), target_qubits, control_qubits)
        setattr(State, name, single)
        setattr(State, f"{name}_multi", multi)
        setattr(State, f"c{name}_multi", cmulti)
    angled = {"p": PhaseShift, "rx": RotateX, "ry": RotateY, "rz": RotateZ}
    for name, cls in angled.items():
        def single(self, index, angle, _c=cls):
            return _c(angle).apply(self, [index], [])

        def multi(self, qubits, angle, _c=cls):
            return self._multi(_c(angle), qubits)

        def cmulti(self, target_qubits, control_qubits, angle, _c=cls):
            return self._multi(_c(angle), target_qubits, control_qubits)
        setattr(State, name, single)
        setattr(State, f"{name}_multi", multi)
        setattr(State, f"c{name}_multi", cmulti)


_install_gate_methods()


def _unitary(self, index, unitary):  # state.rs:1956-1959
    return Unitary2.new(unitary).apply(self, [index], [])


def _unitary_multi(self, qubits, unitary):  # state.rs:1982-1993
    return self._multi(Unitary2.new(unitary), qubits)


def _cunitary_multi(self, target_qubits, control_qubits, unitary):  # state.rs:2018-2030
    return self._multi(Unitary2.new(unitary), target_qubits, control_qubits)


State.unitary, State.unitary_multi, State.cunitary_multi = _unitary, _unitary_multi, _cunitary_multi
State.ry_phase = lambda self, index, angle, phase: Unitary2.from_ry_phase(angle, phase).apply(self, [index], [])
State.ry_phase_multi = lambda self, qubits, angle, phase: self._multi(Unitary2.from_ry_phase(angle, phase), qubits)
State.cry_phase_gates = lambda self, t, c, angle, phase: self._multi(Unitary2.from_ry_phase(angle, phase), t, c)
State.ry_phase_dag = lambda self, q, angle, phase: Unitary2.from_ry_phase_dagger(angle, phase).apply(self, [q], [])
State.ry_phase_dag_multi = lambda self, qubits, angle, phase: self._multi(
    Unitary2.from_ry_phase_dagger(angle, phase), qubits)
State.cry_phase_dag_gates = lambda self, t, c, angle, phase: self._multi(
    Unitary2.from_ry_phase_dagger(angle, phase), t, c)
State.cnot = lambda self, control, target: CNOT().apply(self, [target], [control])  # state.rs:2230
State.swap = lambda self, q1, q2: SWAP().apply(self, [q1, q2], [])
State.cswap = lambda self, t1, t2, controls: SWAP().apply(self, [t1, t2], controls)
State.matchgate = lambda self, target, theta, phi1, phi2: Matchgate(theta, phi1, phi2).apply(self, [target], [])
State.cmatchgate = lambda self, target, theta, phi1, phi2, controls: Matchgate(theta, phi1, phi2).apply(
    self, [target], controls)
State.toffoli = lambda self, c1, c2, target: Toffoli().apply(self, [target], [c1, c2])  # state.rs:2343


# ----------------------------------------------------------------------------
class PauliString:
    """pauli_string.rs:13-287."""

    def __init__(self, coefficient: complex):
        self._ops = {}
        self._coefficient = complex(coefficient)

    @staticmethod
    def new(coefficient):
        return PauliString(coefficient)

    @staticmethod
    def with_ops(coefficient, ops: dict):
        p = PauliString(coefficient)
        p._ops = dict(ops)
        return p

    def __len__(self):
        return len(self._ops)

    def len(self):
        return len(self._ops)

    def coefficient(self) -> complex:
        return self._coefficient

    def ops(self) -> dict:
        return self._ops

    def to_gates(self):  # pauli_string.rs:118-122
        return [Gate.Operator(op, [q], []) for q, op in self._ops.items()]

    def add_op(self, qubit: int, op):
        if qubit in self._ops:
            raise RuntimeError(f"Duplicate Pauli string operator for qubit: {qubit}")  # panic, 66-70
        self._ops[qubit] = op

    def with_op(self, qubit, op):
        self.add_op(qubit, op)
        return self

    def get_targets(self):
        return sorted(self._ops.keys())

    def _arrays(self):
        qs = list(self._ops.keys())
        ps = [{G_X: 1, G_Y: 2, G_Z: 3}[self._ops[q].KIND] for q in qs]
        return qs, ps

    def _check(self, state):
        for q in self._ops:  # each single-Pauli apply validates its own target (operator.rs:481)
            if q >= state.num_qubits:
                raise Error("InvalidQubitIndex", q, state.num_qubits)

    def apply(self, state: State) -> State:  # pauli_string.rs:139-151
        self._check(state)
        qs, ps = self._arrays()
        out = state.clone()
        c = self._coefficient
        _lib.orc_pauli_apply(out._ptr(), state.num_qubits, _u32(qs), (C.c_int32 * max(1, len(ps)))(*ps),
                             len(qs), _dbl([c.real, c.imag]))
        return out

    def apply_operators(self, state: State) -> State:  # pauli_string.rs:172-184
        self._check(state)
        qs, ps = self._arrays()
        out = state.clone()
        _lib.orc_pauli_apply(out._ptr(), state.num_qubits, _u32(qs), (C.c_int32 * max(1, len(ps)))(*ps),
                             len(qs), None)
        return out

    def apply_normalised(self, state: State) -> State:  # pauli_string.rs:165-168
        return self.apply_operators(state).normalise()

    def apply_exp(self, state: State) -> State:  # pauli_string.rs:198-223
        return self._exp(state, self._coefficient)

    def apply_exp_factor(self, state: State, factor: complex) -> State:  # pauli_string.rs:237-262
        return self._exp(state, c_mul(self._coefficient, complex(factor)))

    def _exp(self, state, alpha):
        self._check(state)
        qs, ps = self._arrays()
        out = state.clone()
        _lib.orc_pauli_exp(out._ptr(), state.num_qubits, _u32(qs), (C.c_int32 * max(1, len(ps)))(*ps),
                           len(qs), _dbl([alpha.real, alpha.imag]))
        return out

    def apply_exp_neg_i_dt(self, state: State, dt: float) -> State:  # pauli_string.rs:281-287
        if self._coefficient.imag != 0.0:
            raise Error("InvalidPauliStringCoefficient", self._coefficient)
        return self.apply_exp_factor(state, complex(0.0, -dt))

    def hermitian_conjugate(self):
        return PauliString.with_ops(self._coefficient.conjugate(), self._ops)

    def expect(self, state: State) -> complex:
        self._check(state)
        qs, ps = self._arrays()
        out = (C.c_double * 2)()
        c = self._coefficient
        _lib.orc_pauli_expect(state._ptr(), state.num_qubits, _u32(qs), (C.c_int32 * max(1, len(ps)))(*ps),
                              len(qs), _dbl([c.real, c.imag]), out)
        return complex(out[0], out[1])

    def __mul__(self, rhs):
        return PauliString.with_ops(c_mul(self._coefficient, complex(rhs)), self._ops)

    __rmul__ = __mul__

    def __add__(self, other):
        return SumOp([self, other])


class SumOp:
    """pauli_string.rs:398-507."""

    def __init__(self, terms: Iterable[PauliString]):
        self.terms = list(terms)

    @staticmethod
    def new(terms):
        return SumOp(terms)

    def num_terms(self):
        return len(self.terms)

    def add_term(self, term):
        self.terms.append(term)

    def with_term(self, term):
        self.add_term(term)
        return self

    def apply(self, state: State) -> State:  # pauli_string.rs:453-466
        if not self.terms:
            return state * 0.0
        acc = None
        for t in self.terms:
            s = t.apply(state)
            acc = s if acc is None else acc + s
        return acc

    def expectation_value(self, state: State) -> complex:  # pauli_string.rs:485-507
        total = complex(0.0, 0.0)
        for t in self.terms:
            total = total + t.expect(state)
        return total

    def __mul__(self, rhs):
        return SumOp([t * rhs for t in self.terms])

    def __add__(self, other):
        if isinstance(other, PauliString):
            return SumOp(self.terms + [other])
        return SumOp(self.terms + other.terms)


# ----------------------------------------------------------------------------
class Gate:
    """gate.rs:13-122.  kind in {Operator, Measurement, PauliString, PauliTimeEvolution}.
    (Parametric gates resolve to Operator gates before touching amplitudes, gate.rs:107-114.)"""

    def __init__(self, kind, **kw):
        self.kind = kind
        self.__dict__.update(kw)

    @staticmethod
    def Operator(op, targets, controls=()):
        return Gate("Operator", op=op, targets=list(targets), controls=list(controls))

    @staticmethod
    def Measurement(basis, indices):
        return Gate("Measurement", basis=basis, targets=list(indices), controls=[])

    @staticmethod
    def PauliString(ps):
        return Gate("PauliString", pauli_string=ps, targets=ps.get_targets(), controls=[])

    @staticmethod
    def PauliTimeEvolution(ps, time):
        return Gate("PauliTimeEvolution", pauli_string=ps, time=float(time), targets=ps.get_targets(), controls=[])

    @staticmethod
    def Parametric(p_gate, targets, controls=()):
        return Gate("Parametric", p_gate=p_gate, targets=list(targets), controls=list(controls))

    def concrete(self):  # circuit.rs:205-217
        if self.kind == "Parametric":
            return self.p_gate.to_concrete_gates(self.targets, self.controls)
        if self.kind == "PauliString":
            return self.pauli_string.to_gates()
        return [self]

    def apply(self, state: State, seed: Optional[int] = None) -> State:  # gate.rs:99-122
        if self.kind == "Operator":
            return self.op.apply(state, self.targets, self.controls)
        if self.kind == "Parametric":                                     # gate.rs:107-114
            cur = state.clone()
            for g in self.p_gate.to_concrete_gates(self.targets, self.controls):
                cur = g.apply(cur)
            return cur
        if self.kind == "Measurement":
            return state.measure(self.basis, self.targets, seed=seed).new_state
        if self.kind == "PauliString":
            return self.pauli_string.apply_normalised(state)
        return self.pauli_string.apply_exp_neg_i_dt(state, self.time)

    def get_target_qubits(self):
        return self.targets

    def get_control_qubits(self):
        return self.controls if self.kind in ("Operator", "Parametric") else None


class Parameter:
    """parametric/parameter.rs:13-74: `clone` shares the values (Arc), `deep_clone` copies them."""

    def __init__(self, initial_values, _shared=None):
        self._v = _shared if _shared is not None else {"values": [float(x) for x in initial_values]}

    @staticmethod
    def new(initial_values):
        return Parameter(initial_values)

    def clone(self):
        return Parameter((), _shared=self._v)

    def deep_clone(self):
        return Parameter(self.get())

    def get(self):
        return list(self._v["values"])

    def set(self, new_values):
        if len(new_values) != len(self._v["values"]):
            raise ValueError("wrong number of parameter values")
        self._v["values"] = [float(x) for x in new_values]

    def __len__(self):
        return len(self._v["values"])


class ParametricGate:
    """parametric/parametric_gate.rs: `kind` selects which concrete constructor the values feed."""
    KINDS = {"ry_phase": 2, "ry_phase_dag": 2, "matchgate": 3, "rx": 1, "ry": 1, "rz": 1, "p": 1}

    def __init__(self, kind, parameter):
        assert len(parameter) == self.KINDS[kind]
        self.kind, self.parameter = kind, parameter

    def to_concrete_gates(self, targets, controls):
        v = self.parameter.get()
        if self.kind == "matchgate":                                   # parametric_gate.rs:95-104
            return [Gate.controlled_matchgate(targets[0], list(controls), v[0], v[1], v[2])]
        ctor = getattr(Gate, f"{self.kind}_controlled_gates")          # e.g. Gate::rx_controlled_gates (124-131)
        return ctor(list(targets), list(controls), *v)


def ParametricRyPhase(parameter): return ParametricGate("ry_phase", parameter)            # noqa: E704
def ParametricRyPhaseDag(parameter): return ParametricGate("ry_phase_dag", parameter)     # noqa: E704
def ParametricMatchgate(parameter): return ParametricGate("matchgate", parameter)         # noqa: E704
def ParametricRx(parameter): return ParametricGate("rx", parameter)                       # noqa: E704
def ParametricRy(parameter): return ParametricGate("ry", parameter)                       # noqa: E704
def ParametricRz(parameter): return ParametricGate("rz", parameter)                       # noqa: E704
def ParametricP(parameter): return ParametricGate("p", parameter)                         # noqa: E704


def _install_gate_ctors():
    """gate.rs:165-1267: constructor helpers (same names and argument order)."""
    simple = {"h": Hadamard, "x": lambda: Pauli.X, "y": lambda: Pauli.Y, "z": lambda: Pauli.Z, "i": Identity,
              "s": PhaseS, "s_dag": PhaseSdag, "t": PhaseT, "t_dag": PhaseTdag}
    for name, fac in simple.items():
        setattr(Gate, f"{name}_gate", staticmethod(lambda q, _f=fac: Gate.Operator(_f(), [q], [])))
        setattr(Gate, f"{name}_multi_gate",
                staticmethod(lambda qs, _f=fac: [Gate.Operator(_f(), [q], []) for q in qs]))
        setattr(Gate, f"{name}_controlled_gates",
                staticmethod(lambda ts, cs, _f=fac: [Gate.Operator(_f(), [q], list(cs)) for q in ts]))
    for name, cls in {"p": PhaseShift, "rx": RotateX, "ry": RotateY, "rz": RotateZ}.items():
        setattr(Gate, f"{name}_gate", staticmethod(lambda q, a, _c=cls: Gate.Operator(_c(a), [q], [])))
        setattr(Gate, f"{name}_multi_gate",
                staticmethod(lambda qs, a, _c=cls: [Gate.Operator(_c(a), [q], []) for q in qs]))
        setattr(Gate, f"{name}_controlled_gates",
                staticmethod(lambda ts, cs, a, _c=cls: [Gate.Operator(_c(a), [q], list(cs)) for q in ts]))
    Gate.unitary2_gate = staticmethod(lambda q, u: Gate.Operator(Unitary2.new(u), [q], []))
    Gate.unitary2_multi_gate = staticmethod(lambda qs, u: [Gate.Operator(Unitary2.new(u), [q], []) for q in qs])
    Gate.unitary2_controlled_gates = staticmethod(
        lambda ts, cs, u: [Gate.Operator(Unitary2.new(u), [q], list(cs)) for q in ts])
    Gate.ry_phase_gate = staticmethod(lambda q, th, ph: Gate.Operator(Unitary2.from_ry_phase(th, ph), [q], []))
    Gate.ry_phase_dag_gate = staticmethod(
        lambda q, th, ph: Gate.Operator(Unitary2.from_ry_phase_dagger(th, ph), [q], []))
    Gate.ry_phase_multi_gate = staticmethod(
        lambda qs, th, ph: [Gate.Operator(Unitary2.from_ry_phase(th, ph), [q], []) for q in qs])
    Gate.ry_phase_controlled_gates = staticmethod(
        lambda ts, cs, th, ph: [Gate.Operator(Unitary2.from_ry_phase(th, ph), [q], list(cs)) for q in ts])
    Gate.ry_phase_dag_multi_gate = staticmethod(
        lambda qs, th, ph: [Gate.Operator(Unitary2.from_ry_phase_dagger(th, ph), [q], []) for q in qs])
    Gate.ry_phase_dag_controlled_gates = staticmethod(
        lambda ts, cs, th, ph: [Gate.Operator(Unitary2.from_ry_phase_dagger(th, ph), [q], list(cs)) for q in ts])
    Gate.cnot_gate = staticmethod(lambda target, control: Gate.Operator(CNOT(), [target], [control]))  # gate.rs:1128
    Gate.swap_gate = staticmethod(lambda q1, q2: Gate.Operator(SWAP(), [q1, q2], []))
    Gate.swap_controlled_gate = staticmethod(lambda q1, q2, cs: Gate.Operator(SWAP(), [q1, q2], list(cs)))
    Gate.toffoli_gate = staticmethod(lambda target, controls: Gate.Operator(Toffoli(), [target], list(controls)))
    Gate.pauli_string_gate = staticmethod(lambda ps: Gate.PauliString(ps))
    Gate.pauli_time_evolution_gate = staticmethod(lambda ps, t: Gate.PauliTimeEvolution(ps, t))
    Gate.matchgate = staticmethod(lambda t, th, p1, p2: Gate.Operator(Matchgate(th, p1, p2), [t], []))
    Gate.controlled_matchgate = staticmethod(
        lambda t, cs, th, p1, p2: Gate.Operator(Matchgate(th, p1, p2), [t], list(cs)))


class Subroutine:
    """subroutine.rs:13-160."""

    def __init__(self, gates, num_qubits):
        self.gates, self.num_qubits = list(gates), num_qubits

    @staticmethod
    def qft(qubits: Sequence[int], num_qubits: int) -> "Subroutine":  # subroutine.rs:90-112
        b = CircuitBuilder(num_qubits)
        n = len(qubits)
        for i in range(n):
            b.h_gate(qubits[i])
            den = 2.0
            for k in range(1, n - i):
                b.cp_gates([qubits[i]], [qubits[i + k]], math.pi / den)
                den *= 2.0
        for i in range(n // 2):
            b.swap_gate(qubits[i], qubits[n - 1 - i])
        return b.build_subroutine()

    @staticmethod
    def iqft(qubits: Sequence[int], num_qubits: int) -> "Subroutine":  # subroutine.rs:125-160
        b = CircuitBuilder(num_qubits)
        n = len(qubits)
        for i in range(n // 2):
            b.swap_gate(qubits[i], qubits[n - 1 - i])
        for i in reversed(range(n)):
            if n > i + 1:
                k_initial = (n - 1) - i
                den = 2.0 ** k_initial
                for it in range(k_initial):
                    k = k_initial - it
                    b.cp_gates([qubits[i]], [qubits[i + k]], -math.pi / den)
                    if k > 1:
                        den /= 2.0
            b.h_gate(qubits[i])
        return b.build_subroutine()


class Circuit:
    """circuit.rs:27-202."""

    def __init__(self, num_qubits: int):
        self.gates: List[Gate] = []
        self.num_qubits = num_qubits

    @staticmethod
    def _validate(gate: Gate, n: int):  # circuit.rs:35-52
        for q in gate.get_target_qubits():
            if q >= n:
                raise Error("InvalidQubitIndex", q, n)
        for q in gate.get_control_qubits() or []:
            if q >= n:
                raise Error("InvalidQubitIndex", q, n)

    @staticmethod
    def with_gates(gates, num_qubits):
        c = Circuit(num_qubits)
        for g in gates:
            Circuit._validate(g, num_qubits)
        c.gates = list(gates)
        return c

    def add_gate(self, gate):
        Circuit._validate(gate, self.num_qubits)
        self.gates.append(gate)

    def add_gates(self, gates):
        for g in gates:
            Circuit._validate(g, self.num_qubits)
        self.gates.extend(gates)

    def get_num_qubits(self):
        return self.num_qubits

    def get_gates(self):
        return self.gates

    def execute(self, initial_state: State, seed: Optional[int] = None) -> State:  # circuit.rs:160-172
        if initial_state.num_qubits != self.num_qubits:
            raise Error("InvalidNumberOfQubits", initial_state.num_qubits)
        cur = initial_state.clone()
        for i, g in enumerate(self.gates):
            cur = g.apply(cur, None if seed is None else seed + i)
        return cur

    def to_concrete_circuit(self):  # circuit.rs:204-221
        c = Circuit(self.num_qubits)
        c.gates = [cg for g in self.gates for cg in g.concrete()]
        return c

    def trace_execution(self, initial_state: State) -> List[State]:  # circuit.rs:188-202
        if initial_state.num_qubits != self.num_qubits:
            raise Error("InvalidNumberOfQubits", initial_state.num_qubits)
        cur = initial_state.clone()
        out = [cur.clone()]
        for g in self.gates:
            cur = g.apply(cur)
            out.append(cur.clone())
        return out


class CircuitBuilder:
    """circuit.rs:288-1742 (operator, Pauli and measurement adders; parametric adders are
    host-only sugar that resolve to these)."""

    def __init__(self, num_qubits: int):
        self.gates: List[Gate] = []
        self.num_qubits = num_qubits

    @staticmethod
    def new(num_qubits: int):  # circuit.rs:296-301
        return CircuitBuilder(num_qubits)

    def add_gate(self, gate):
        self.gates.append(gate)
        return self

    def add_gates(self, gates):
        self.gates.extend(gates)
        return self

    def build(self) -> Circuit:  # circuit.rs:340-343
        c = Circuit.with_gates(self.gates, self.num_qubits)
        return c

    def build_final(self) -> Circuit:
        c = Circuit.with_gates(self.gates, self.num_qubits)
        self.gates = []
        return c

    def build_subroutine(self) -> Subroutine:
        s = Subroutine(self.gates, self.num_qubits)
        self.gates = []
        return s

    def add_subroutine(self, sub: Subroutine):  # circuit.rs:373-376
        self.gates.extend(sub.gates)
        return self

    def _each(self, op, targets, controls=()):
        for q in targets:
            self.gates.append(Gate.Operator(op, [q], controls))
        return self

    def cnot_gate(self, target_qubit, control_qubit):  # circuit.rs:1071 (target first!)
        return self.add_gate(Gate.Operator(CNOT(), [target_qubit], [control_qubit]))

    def swap_gate(self, q1, q2):
        return self.add_gate(Gate.Operator(SWAP(), [q1, q2], []))

    def cswap_gate(self, t1, t2, controls):
        return self.add_gate(Gate.Operator(SWAP(), [t1, t2], controls))

    def toffoli_gate(self, control1, control2, target):  # circuit.rs:1118-1123
        return self.add_gate(Gate.Operator(Toffoli(), [target], [control1, control2]))

    def pauli_string_gate(self, ps):
        return self.add_gate(Gate.PauliString(ps))

    def pauli_time_evolution_gate(self, ps, time):
        return self.add_gate(Gate.PauliTimeEvolution(ps, time))

    def matchgate(self, target, theta, phi1, phi2):
        return self.add_gate(Gate.Operator(Matchgate(theta, phi1, phi2), [target], []))

    def cmatchgate(self, target, controls, theta, phi1, phi2):  # circuit.rs:1194-1200: controls SECOND (State::cmatchgate has them last)
        return self.add_gate(Gate.Operator(Matchgate(theta, phi1, phi2), [target], list(controls)))

    def add_operator_gate(self, op, targets, controls=()):
        return self.add_gate(Gate.Operator(op, targets, controls))

    def unitary_gate(self, qubit, unitary):
        return self.add_gate(Gate.Operator(Unitary2.new(unitary), [qubit], []))

    def unitary_gates(self, qubits, unitary):
        return self._each(Unitary2.new(unitary), qubits)

    def cunitary_gates(self, targets, controls, unitary):
        return self._each(Unitary2.new(unitary), targets, controls)

    def ry_phase_gate(self, qubit, theta, phi):
        return self.add_gate(Gate.Operator(Unitary2.from_ry_phase(theta, phi), [qubit], []))

    def ry_phase_gates(self, qubits, theta, phi):
        return self._each(Unitary2.from_ry_phase(theta, phi), qubits)

    def cry_phase_gates(self, targets, controls, theta, phi):
        return self._each(Unitary2.from_ry_phase(theta, phi), targets, controls)

    def ry_phase_dag_gate(self, qubit, theta, phi):
        return self.add_gate(Gate.Operator(Unitary2.from_ry_phase_dagger(theta, phi), [qubit], []))

    def ry_phase_dag_gates(self, qubits, theta, phi):
        return self._each(Unitary2.from_ry_phase_dagger(theta, phi), qubits)

    def cry_phase_dag_gates(self, targets, controls, theta, phi):
        return self._each(Unitary2.from_ry_phase_dagger(theta, phi), targets, controls)

    def measure_gate(self, basis, qubits):
        return self.add_gate(Gate.Measurement(basis, qubits))


def _install_builder_methods():
    simple = {"h": Hadamard, "x": lambda: Pauli.X, "y": lambda: Pauli.Y, "z": lambda: Pauli.Z,
              "s": PhaseS, "t": PhaseT, "sdag": PhaseSdag, "tdag": PhaseTdag}
    for name, fac in simple.items():
        setattr(CircuitBuilder, f"{name}_gate", lambda self, q, _f=fac: self._each(_f(), [q]))
        setattr(CircuitBuilder, f"{name}_gates", lambda self, qs, _f=fac: self._each(_f(), qs))
        setattr(CircuitBuilder, f"c{name}_gates", lambda self, t, c, _f=fac: self._each(_f(), t, c))
    CircuitBuilder.id_gate = lambda self, q: self._each(Identity(), [q])
    CircuitBuilder.id_gates = lambda self, qs: self._each(Identity(), qs)
    CircuitBuilder.ci_gates = lambda self, t, c: self._each(Identity(), t, c)
    for name, cls in {"p": PhaseShift, "rx": RotateX, "ry": RotateY, "rz": RotateZ}.items():
        setattr(CircuitBuilder, f"{name}_gate", lambda self, q, a, _c=cls: self._each(_c(a), [q]))
        setattr(CircuitBuilder, f"{name}_gates", lambda self, qs, a, _c=cls: self._each(_c(a), qs))
        setattr(CircuitBuilder, f"c{name}_gates", lambda self, t, c, a, _c=cls: self._each(_c(a), t, c))

    # parametric adders (circuit.rs:1226-1742)
    def each(self, kind, targets, controls, parameters):
        targets, parameters = list(targets), list(parameters)
        if len(targets) != len(parameters):
            raise Error("MismatchedNumberOfParameters", len(targets), len(parameters))
        for t, prm in zip(targets, parameters):
            self.add_gate(Gate.Parametric(ParametricGate(kind, prm), [t], list(controls)))
        return self
    for kind in ("ry_phase", "ry_phase_dag", "rx", "ry", "rz", "p"):
        setattr(CircuitBuilder, f"parametric_{kind}_gate",
                lambda self, t, prm, _k=kind: self.add_gate(Gate.Parametric(ParametricGate(_k, prm), [t], [])))
        setattr(CircuitBuilder, f"parametric_{kind}_gates", lambda self, ts, prms, _k=kind: each(self, _k, ts, [], prms))
        setattr(CircuitBuilder, f"parametric_c{kind}_gates", lambda self, ts, cs, prms, _k=kind: each(self, _k, ts, cs, prms))
    CircuitBuilder.parametric_matchgate = lambda self, t, prm: self.add_gate(
        Gate.Parametric(ParametricGate("matchgate", prm), [t], []))
    CircuitBuilder.parametric_cmatchgate = lambda self, t, cs, prm: self.add_gate(
        Gate.Parametric(ParametricGate("matchgate", prm), [t], list(cs)))


_install_builder_methods()
_install_gate_ctors()


# ----------------------------------------------------------------------------
class TrotterOrder:
    First = "First"
    Second = "Second"


def first_order_trotter_step(h: SumOp, state: State, dt: float) -> State:  # time_evolution.rs:45-66
    if h.num_terms() == 0:
        raise Error("InvalidNumberOfQubits", 0)
    cur = state.clone()
    for t in h.terms:
        cur = t.apply_exp_factor(cur, complex(0.0, -dt))
    return cur


def second_order_trotter_step(h: SumOp, state: State, dt: float) -> State:  # time_evolution.rs:89-115
    if h.num_terms() == 0:
        raise Error("InvalidNumberOfQubits", 0)
    half = complex(0.0, -dt / 2.0)
    cur = state.clone()
    for t in h.terms:
        cur = t.apply_exp_factor(cur, half)
    for t in reversed(h.terms):
        cur = t.apply_exp_factor(cur, half)
    return cur


def trotter_evolve_state(h: SumOp, state: State, dt: float, num_steps: int, order) -> State:  # 140-167
    if h.num_terms() == 0:
        raise Error("InvalidNumberOfQubits", 0)
    cur = state.clone()
    for _ in range(num_steps):
        cur = first_order_trotter_step(h, cur, dt) if order == TrotterOrder.First \
            else second_order_trotter_step(h, cur, dt)
    return cur


def heisenberg_1d(n: int, jx: float, jy: float, jz: float, h: float, mu: float) -> SumOp:
    """models/heisenberg.rs:28-102.  Term order per site: XX, YY, ZZ, Z(i); periodic neighbour;
    coefficients -J/2 and (code is ground truth, heisenberg.rs:49) -mu * (-h/2) = +mu*h/2."""
    if n < 2:
        raise Error("InvalidNumberOfInputs", n, 2)
    if jx == 0.0 and jy == 0.0 and jz == 0.0 and h == 0.0:
        return SumOp([])
    cx, cy, cz = complex(-0.5 * jx, 0.0), complex(-0.5 * jy, 0.0), complex(-0.5 * jz, 0.0)
    field = complex(-mu * (-0.5 * h), -mu * 0.0)
    terms = []
    for i in range(n):
        j = (i + 1) % n
        if jx != 0.0:
            terms.append(PauliString(cx).with_op(i, Pauli.X).with_op(j, Pauli.X))
        if jy != 0.0:
            terms.append(PauliString(cy).with_op(i, Pauli.Y).with_op(j, Pauli.Y))
        if jz != 0.0:
            terms.append(PauliString(cz).with_op(i, Pauli.Z).with_op(j, Pauli.Z))
        if h != 0.0:
            terms.append(PauliString(field).with_op(i, Pauli.Z))
    return SumOp(terms)


def heisenberg_2d(n_rows: int, m_cols: int, jx: float, jy: float, jz: float, h_field: float, mu: float) -> SumOp:
    """models/heisenberg.rs:122-220 restated literally: for every site index (row-major) push the field term, the
    three vertical-bond terms and the three horizontal-bond terms; field coefficient mu * (-0.5 h) (line 143)."""
    if n_rows < 2:
        raise Error("InvalidNumberOfInputs", n_rows, 2)
    elif m_cols < 2:
        raise Error("InvalidNumberOfInputs", m_cols, 2)
    if jx == 0.0 and jy == 0.0 and jz == 0.0 and h_field == 0.0:
        return SumOp([])
    coeff_x, coeff_y, coeff_z = complex(-0.5 * jx, 0.0), complex(-0.5 * jy, 0.0), complex(-0.5 * jz, 0.0)
    field_coeff = complex(mu * (-0.5 * h_field), mu * 0.0)
    terms = []
    for site in range(n_rows * m_cols):
        r, c = site // m_cols, site % m_cols
        if h_field != 0.0:
            terms.append(PauliString(field_coeff).with_op(site, Pauli.Z))
        down = ((r + 1) % n_rows) * m_cols + c
        if jx != 0.0:
            terms.append(PauliString(coeff_x).with_op(site, Pauli.X).with_op(down, Pauli.X))
        if jy != 0.0:
            terms.append(PauliString(coeff_y).with_op(site, Pauli.Y).with_op(down, Pauli.Y))
        if jz != 0.0:
            terms.append(PauliString(coeff_z).with_op(site, Pauli.Z).with_op(down, Pauli.Z))
        right = r * m_cols + ((c + 1) % m_cols)
        if jx != 0.0:
            terms.append(PauliString(coeff_x).with_op(site, Pauli.X).with_op(right, Pauli.X))
        if jy != 0.0:
            terms.append(PauliString(coeff_y).with_op(site, Pauli.Y).with_op(right, Pauli.Y))
        if jz != 0.0:
            terms.append(PauliString(coeff_z).with_op(site, Pauli.Z).with_op(right, Pauli.Z))
    return SumOp(terms)


def ising_1d(h, j, mu: float) -> SumOp:
    """models/ising.rs:27-75."""
    n = len(h)
    if len(j) != n:
        raise Error("MismatchedNumberOfParameters", n, len(j))
    if n < 2:
        raise Error("InvalidNumberOfInputs", n, 2)
    if all(v == 0.0 for v in h) and all(v == 0.0 for v in j):
        return SumOp([])
    terms = []
    for i in range(n):
        if j[i] != 0.0:
            terms.append(PauliString(complex(j[i], 0.0) * -1.0).with_op(i, Pauli.Z).with_op((i + 1) % n, Pauli.Z))
        if h[i] != 0.0:
            terms.append(PauliString(-1.0 * mu * complex(h[i], 0.0)).with_op(i, Pauli.Z))
    return SumOp(terms)


def ising_1d_uniform(n: int, h: float, j: float, mu: float) -> SumOp:
    """models/ising.rs:90-139."""
    if n < 2:
        raise Error("InvalidNumberOfInputs", n, 2)
    if h == 0.0 and j == 0.0:
        return SumOp([])
    terms = []
    for i in range(n):
        if j != 0.0:
            terms.append(PauliString(complex(j, 0.0) * -1.0).with_op(i, Pauli.Z).with_op((i + 1) % n, Pauli.Z))
        if h != 0.0:
            terms.append(PauliString(-1.0 * mu * complex(h, 0.0)).with_op(i, Pauli.Z))
    return SumOp(terms)


def ising_2d(h, j, mu: float) -> SumOp:
    """models/ising.rs:161-244."""
    n = len(h)
    m = len(h[0]) if n else 0
    if n < 2:
        raise Error("InvalidNumberOfInputs", n, 2)
    elif m < 2:
        raise Error("InvalidNumberOfInputs", m, 2)
    if all(h[r][c] == 0.0 and j[r][c][0] == 0.0 and j[r][c][1] == 0.0 for r in range(n) for c in range(m)):
        return SumOp([])
    terms = []
    for idx in range(n * m):
        r, c = idx // m, idx % m
        if h[r][c] != 0.0:
            terms.append(PauliString(-1.0 * mu * complex(h[r][c], 0.0)).with_op(idx, Pauli.Z))
        if j[r][c][0] != 0.0:
            terms.append(PauliString(complex(j[r][c][0], 0.0) * -1.0).with_op(idx, Pauli.Z).with_op(((r + 1) % n) * m + c, Pauli.Z))
        if j[r][c][1] != 0.0:
            terms.append(PauliString(complex(j[r][c][1], 0.0) * -1.0).with_op(idx, Pauli.Z).with_op(r * m + ((c + 1) % m), Pauli.Z))
    return SumOp(terms)


def ising_2d_uniform(n: int, m: int, h: float, j: float, mu: float) -> SumOp:
    """models/ising.rs:259-324."""
    if n < 2:
        raise Error("InvalidNumberOfInputs", n, 2)
    elif m < 2:
        raise Error("InvalidNumberOfInputs", m, 2)
    if h == 0.0 and j == 0.0:
        return SumOp([])
    terms = []
    for idx in range(n * m):
        r, c = idx // m, idx % m
        if h != 0.0:
            terms.append(PauliString(-1.0 * mu * complex(h, 0.0)).with_op(idx, Pauli.Z))
        if j != 0.0:
            terms.append(PauliString(complex(j, 0.0) * -1.0).with_op(idx, Pauli.Z).with_op(((r + 1) % n) * m + c, Pauli.Z))
            terms.append(PauliString(complex(j, 0.0) * -1.0).with_op(idx, Pauli.Z).with_op(r * m + ((c + 1) % m), Pauli.Z))
    return SumOp(terms)


# ----------------------------------------------------------------------------
# Synthetic workloads shared by oracle, CPU baseline and GPU engine
def random_state(num_qubits: int, seed: int = 20260002) -> State:
    v = np.empty(1 << num_qubits, dtype=np.complex128)
    _lib.orc_random_state(v.ctypes.data_as(C.c_void_p), num_qubits, C.c_uint64(seed))
    return State(v, num_qubits)


def gate_faithful(src: np.ndarray, dst: np.ndarray, n: int, kind: int, targets, controls, params):
    """One gate by the reference's rayon-branch pass structure (timed CPU baseline)."""
    _lib.orc_gate_faithful(src.ctypes.data_as(C.c_void_p), dst.ctypes.data_as(C.c_void_p), n, kind,
                           _u32(targets), len(targets), _u32(controls), len(controls), _dbl(params))


def gate_inplace(v: np.ndarray, n: int, kind: int, targets, controls, params):
    _lib.orc_gate(v.ctypes.data_as(C.c_void_p), n, kind, _u32(targets), len(targets),
                  _u32(controls), len(controls), _dbl(params))
