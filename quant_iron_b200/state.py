"""`State` (src/components/state.rs) over a device-resident Complex<f64> amplitude buffer.

The reference's methods take `&self` and return a new `State`; here that is a device-to-device
clone followed by the in-place kernel.  Every method also has an in-place twin with a trailing
underscore (`h_`, `apply_`, `measure_`) for states that do not fit in HBM twice (33 qubits).
`state_vector` is a property that copies the amplitudes to the host: the reference's public
`Vec<Complex<f64>>` field cannot exist for a device state (documented API break).
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import List, Optional, Sequence

import numpy as np

from . import _ffi
from .errors import Error
from .measurement import MeasurementBasis, MeasurementResult
from .operators import (CNOT, SWAP, Hadamard, Identity, Matchgate, Operator, Pauli, PhaseS, PhaseSdag,
                        PhaseShift, PhaseT, PhaseTdag, RotateX, RotateY, RotateZ, Toffoli, Unitary2)

F32_EPS = 1.1920928955078125e-07   # State::eq tolerance, state.rs:2360-2367
_lib = _ffi.lib


def _new_handle(fn, *args) -> C.c_void_p:
    h = C.c_void_p()
    _ffi.check(fn(*args, C.byref(h)))
    return h


class State:
    """state.rs:74-81."""

    def __init__(self, state_vector=None, num_qubits: int = 0, *, _handle=None):
        # State { state_vector, num_qubits } literal: no checks (state_tests.rs:145-150)
        if _handle is not None:
            self._h = _handle
        else:
            v = np.ascontiguousarray(state_vector, dtype=np.complex128)
            self._h = _new_handle(_lib.qi_state_from_host, v.ctypes.data_as(C.c_void_p), v.shape[0],
                                  int(num_qubits), 0)
        self.num_qubits = int(_lib.qi_state_num_qubits(self._h))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and _lib is not None:
            try:
                _lib.qi_state_free(h)
            except Exception:
                pass
            self._h = None

    # ---- host access ----
    def __len__(self):
        return int(_lib.qi_state_len(self._h))

    @property
    def state_vector(self) -> np.ndarray:
        out = np.empty(len(self), dtype=np.complex128)
        _ffi.check(_lib.qi_state_to_host(self._h, out.ctypes.data_as(C.c_void_p), out.shape[0]))
        return out

    def to_host(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        if out is None:
            return self.state_vector
        _ffi.check(_lib.qi_state_to_host(self._h, out.ctypes.data_as(C.c_void_p), out.shape[0]))
        return out

    def upload_(self, state_vector: np.ndarray) -> "State":
        """Overwrite the amplitudes (this rank's shard for a sharded state) from a host array."""
        v = np.ascontiguousarray(state_vector, dtype=np.complex128)
        _ffi.check(_lib.qi_state_upload(self._h, v.ctypes.data_as(C.c_void_p), v.shape[0]))
        return self

    def clone(self) -> "State":
        return State(_handle=_new_handle(_lib.qi_state_clone, self._h))

    # ---- constructors ----
    @staticmethod
    def new(state_vector) -> "State":  # state.rs:99-127
        v = np.ascontiguousarray(state_vector, dtype=np.complex128)
        return State(_handle=_new_handle(_lib.qi_state_from_host, v.ctypes.data_as(C.c_void_p), v.shape[0], 0, 1))

    @staticmethod
    def from_host(state_vector, check: bool = True) -> "State":
        v = np.ascontiguousarray(state_vector, dtype=np.complex128)
        n = max(0, v.shape[0].bit_length() - 1)
        return State(_handle=_new_handle(_lib.qi_state_from_host, v.ctypes.data_as(C.c_void_p), v.shape[0], n,
                                         1 if check else 0))

    @staticmethod
    def new_hartree_fock(num_electrons: int, num_orbitals: int) -> "State":  # state.rs:140-151
        if num_orbitals == 0 or num_orbitals < num_electrons:
            raise Error("InvalidInputValue", num_orbitals)
        n = ((1 << num_electrons) - 1) << (num_orbitals - num_electrons)
        return State.new_basis_n(num_orbitals, n)

    @staticmethod
    def new_zero(num_qubits: int) -> "State":
        return State(_handle=_new_handle(_lib.qi_state_new_zero, num_qubits))

    @staticmethod
    def new_basis_n(num_qubits: int, n: int) -> "State":
        return State(_handle=_new_handle(_lib.qi_state_new_basis_n, num_qubits, n))

    @staticmethod
    def new_plus(num_qubits: int) -> "State":
        return State(_handle=_new_handle(_lib.qi_state_new_plus, num_qubits))

    @staticmethod
    def new_minus(num_qubits: int) -> "State":
        return State(_handle=_new_handle(_lib.qi_state_new_minus, num_qubits))

    @staticmethod
    def new_ghz(num_qubits: int) -> "State":
        return State(_handle=_new_handle(_lib.qi_state_new_ghz, num_qubits))

    @staticmethod
    def new_random(num_qubits: int, seed: int = 20260002) -> "State":
        """Synthetic normalised state generated on the device."""
        s = State.new_zero(num_qubits)
        _ffi.check(_lib.qi_state_init_random(s._h, seed))
        return s

    @staticmethod
    def _bell(v) -> "State":  # state.rs:28-66 bell_vectors (FRAC_1_SQRT_2)
        return State(np.array(v, dtype=np.complex128), 2)

    @staticmethod
    def new_phi_plus():
        a = math.sqrt(0.5)
        return State._bell([a, 0, 0, a])

    @staticmethod
    def new_phi_minus():
        a = math.sqrt(0.5)
        return State._bell([a, 0, 0, -a])

    @staticmethod
    def new_psi_plus():
        a = math.sqrt(0.5)
        return State._bell([0, a, a, 0])

    @staticmethod
    def new_psi_minus():
        a = math.sqrt(0.5)
        return State._bell([0, a, -a, 0])

    # ---- accessors ----
    def equals_without_phase(self, other: "State") -> bool:  # state.rs:384-390
        if self.num_qubits != other.num_qubits:
            return False
        return abs(abs(self.inner_product(other)) - 1.0) < F32_EPS

    def conj(self) -> "State":
        out = self.clone()
        _ffi.check(_lib.qi_conj(out._h))
        return out

    def amplitude(self, n: int) -> complex:  # state.rs:448-453
        out = (C.c_double * 2)()
        _ffi.check(_lib.qi_state_amplitude(self._h, n, out))
        return complex(out[0], out[1])

    def probability(self, n: int) -> float:  # state.rs:418-424
        z = self.amplitude(n)
        return z.real * z.real + z.imag * z.imag

    def fs_dist(self, other: "State") -> float:  # state.rs:470-480
        return math.acos(abs(self.normalise().inner_product(other.normalise())))

    def fs_fidelity(self, other: "State") -> float:  # state.rs:492-498
        return abs(self.normalise().inner_product(other.normalise())) ** 2

    def __eq__(self, other):  # state.rs:2348-2372
        if not isinstance(other, State) or self.num_qubits != other.num_qubits or len(self) != len(other):
            return False
        d = self.state_vector - other.state_vector
        return bool(np.all(np.abs(d.real) <= F32_EPS) and np.all(np.abs(d.imag) <= F32_EPS))

    __hash__ = None

    # ---- linear algebra ----
    def inner_product(self, other: "State") -> complex:
        out = (C.c_double * 2)()
        _ffi.check(_lib.qi_inner_product(self._h, other._h, out))
        return complex(out[0], out[1])

    def norm_sqr(self) -> float:
        out = C.c_double()
        _ffi.check(_lib.qi_norm_sqr(self._h, C.byref(out)))
        return float(out.value)

    def normalise(self) -> "State":
        out = self.clone()
        _ffi.check(_lib.qi_normalise(out._h))
        return out

    def normalise_(self) -> "State":
        _ffi.check(_lib.qi_normalise(self._h))
        return self

    def tensor_product(self, other: "State") -> "State":
        return State(_handle=_new_handle(_lib.qi_tensor_product, self._h, other._h))

    def __mul__(self, rhs):
        z = complex(rhs)
        out = self.clone()
        _ffi.check(_lib.qi_scale(out._h, _ffi.dbl_array([z.real, z.imag])))
        return out

    __rmul__ = __mul__

    def __add__(self, rhs: "State"):
        if self.num_qubits != rhs.num_qubits:
            raise RuntimeError("Cannot add states with different numbers of qubits")  # panic, state.rs:2780
        out = self.clone()
        _ffi.check(_lib.qi_add(out._h, rhs._h))
        return out

    def __sub__(self, rhs: "State"):
        if self.num_qubits != rhs.num_qubits:
            raise RuntimeError("Cannot subtract states with different numbers of qubits")
        out = self.clone()
        _ffi.check(_lib.qi_sub(out._h, rhs._h))
        return out

    # ---- operators ----
    def apply_(self, op: Operator, target_qubits: Sequence[int], control_qubits: Sequence[int] = ()) -> "State":
        """Operator::apply in place (no clone)."""
        rec, keep = op.record(list(target_qubits), list(control_qubits))
        _ffi.check(_lib.qi_apply_gate(self._h, C.byref(rec)))
        del keep
        return self

    def operate(self, unitary: Operator, target_qubits, control_qubits=()):  # state.rs:970-1002
        nt, nc = len(target_qubits), len(control_qubits)
        if unitary.base_qubits() != nt + nc:
            raise Error("InvalidNumberOfQubits", unitary.base_qubits())
        if nt > self.num_qubits:
            raise Error("InvalidNumberOfQubits", self.num_qubits)
        for q in list(target_qubits) + list(control_qubits):
            if q >= self.num_qubits:
                raise Error("InvalidQubitIndex", q, self.num_qubits)
        return unitary.apply(self, target_qubits, control_qubits)

    def _multi(self, op: Operator, targets, controls=()):
        out = self.clone()
        for q in targets:
            out.apply_(op, [q], controls)
        return out

    def _multi_(self, op: Operator, targets, controls=()):
        for q in targets:
            self.apply_(op, [q], controls)
        return self

    # ---- measurement (state.rs:525-784) ----
    def _qubits_arg(self, measured_qubits):
        qs = [int(q) for q in measured_qubits]
        return _ffi.u32_array(qs), len(qs)

    def probabilities(self, qubits: Sequence[int]) -> np.ndarray:
        arr, m = self._qubits_arg(qubits)
        nq = m if m else self.num_qubits
        out = np.zeros(1 << min(nq, 40), dtype=np.float64)
        _ffi.check(_lib.qi_probabilities(self._h, arr, m, out.ctypes.data_as(_ffi.dp)))
        return out

    def measure_(self, basis: MeasurementBasis, measured_qubits=(), seed: Optional[int] = None, draw: int = 0):
        """In-place measure: this state becomes the collapsed state; returns (indices, outcomes)."""
        arr, m = self._qubits_arg(measured_qubits)
        if seed is None:
            seed = int.from_bytes(os.urandom(8), "little")
        nq = m if m else self.num_qubits
        outcomes = (C.c_uint8 * max(1, nq))()
        binv = C.c_uint64()
        cu = _ffi.dbl_array(basis.flat_matrix()) if basis.name == "Custom" else None
        _ffi.check(_lib.qi_measure(self._h, basis.code, cu, arr, m, C.c_uint64(seed & (2**64 - 1)), draw, outcomes,
                                   C.byref(binv)))
        indices = list(measured_qubits) if m else list(range(self.num_qubits))
        return indices, [int(outcomes[j]) for j in range(nq)]

    def measure(self, basis: MeasurementBasis, measured_qubits=(), seed: Optional[int] = None) -> MeasurementResult:
        """state.rs:525-730.  The draw is u_0 of the shared-seed stream (unseeded: OS entropy)."""
        new_state = self.clone()
        indices, outcomes = new_state.measure_(basis, measured_qubits, seed, 0)
        return MeasurementResult(basis, indices, outcomes, new_state)

    def measure_n(self, basis: MeasurementBasis, measured_qubits, n: int, seed: Optional[int] = None) -> List[MeasurementResult]:
        """state.rs:750-784: n independent measurements of the same state; shot k draws u_k."""
        if n == 0:
            raise Error("InvalidNumberOfMeasurements", 0)
        if seed is None:
            seed = int.from_bytes(os.urandom(8), "little")
        out = []
        for k in range(n):
            st = self.clone()
            indices, outcomes = st.measure_(basis, measured_qubits, seed, k)
            out.append(MeasurementResult(basis, indices, outcomes, st))
        return out

    def sample_counts(self, measured_qubits, shots: int, seed: int) -> np.ndarray:
        """Outcome bins of `shots` computational-basis measurements (measure_n's outcomes without
        materialising n collapsed states): one probability sweep, device prefix scan, seeded draws."""
        arr, m = self._qubits_arg(measured_qubits)
        bins = np.zeros(max(1, shots), dtype=np.uint64)
        _ffi.check(_lib.qi_sample(self._h, arr, m, shots, C.c_uint64(seed & (2**64 - 1)),
                                  bins.ctypes.data_as(_ffi.u64p)))
        return bins.astype(np.int64)

    def collapse_(self, measured_qubits, bin_value: int) -> "State":
        arr, m = self._qubits_arg(measured_qubits)
        _ffi.check(_lib.qi_collapse(self._h, arr, m, bin_value))
        return self


def _install_gate_methods():
    """state.rs:1019-2345: the convenience methods, generated from a table, each with an
    in-place twin (`h_`, `h_multi_`, `ch_multi_`)."""
    simple = {"h": Hadamard, "x": lambda: Pauli.X, "y": lambda: Pauli.Y, "z": lambda: Pauli.Z, "i": Identity,
              "s": PhaseS, "t": PhaseT, "s_dag": PhaseSdag, "t_dag": PhaseTdag}
    for name, fac in simple.items():
        setattr(State, name, lambda self, index, _f=fac: _f().apply(self, [index], []))
        setattr(State, f"{name}_multi", lambda self, qubits, _f=fac: self._multi(_f(), qubits))
        setattr(State, f"c{name}_multi", lambda self, t, c, _f=fac: self._multi(_f(), t, c))
        setattr(State, f"{name}_", lambda self, index, _f=fac: self.apply_(_f(), [index], []))
        setattr(State, f"{name}_multi_", lambda self, qubits, _f=fac: self._multi_(_f(), qubits))
        setattr(State, f"c{name}_multi_", lambda self, t, c, _f=fac: self._multi_(_f(), t, c))
    for name, cls in {"p": PhaseShift, "rx": RotateX, "ry": RotateY, "rz": RotateZ}.items():
        setattr(State, name, lambda self, index, angle, _c=cls: _c(angle).apply(self, [index], []))
        setattr(State, f"{name}_multi", lambda self, qubits, angle, _c=cls: self._multi(_c(angle), qubits))
        setattr(State, f"c{name}_multi", lambda self, t, c, angle, _c=cls: self._multi(_c(angle), t, c))
        setattr(State, f"{name}_", lambda self, index, angle, _c=cls: self.apply_(_c(angle), [index], []))
        setattr(State, f"{name}_multi_", lambda self, qubits, angle, _c=cls: self._multi_(_c(angle), qubits))
        setattr(State, f"c{name}_multi_", lambda self, t, c, angle, _c=cls: self._multi_(_c(angle), t, c))


_install_gate_methods()

State.unitary = lambda self, index, u: Unitary2.new(u).apply(self, [index], [])                    # state.rs:1956
State.unitary_multi = lambda self, qubits, u: self._multi(Unitary2.new(u), qubits)                  # state.rs:1982
State.cunitary_multi = lambda self, t, c, u: self._multi(Unitary2.new(u), t, c)                     # state.rs:2018
State.ry_phase = lambda self, index, angle, phase: Unitary2.from_ry_phase(angle, phase).apply(self, [index], [])
State.ry_phase_multi = lambda self, qubits, angle, phase: self._multi(Unitary2.from_ry_phase(angle, phase), qubits)
State.cry_phase_gates = lambda self, t, c, angle, phase: self._multi(Unitary2.from_ry_phase(angle, phase), t, c)
State.ry_phase_dag = lambda self, q, angle, phase: Unitary2.from_ry_phase_dagger(angle, phase).apply(self, [q], [])
State.ry_phase_dag_multi = lambda self, qubits, angle, phase: self._multi(
    Unitary2.from_ry_phase_dagger(angle, phase), qubits)
State.cry_phase_dag_gates = lambda self, t, c, angle, phase: self._multi(
    Unitary2.from_ry_phase_dagger(angle, phase), t, c)
State.cnot = lambda self, control, target: CNOT().apply(self, [target], [control])                  # state.rs:2230
State.swap = lambda self, q1, q2: SWAP().apply(self, [q1, q2], [])                                  # state.rs:2248
State.cswap = lambda self, t1, t2, controls: SWAP().apply(self, [t1, t2], controls)                 # state.rs:2268
State.matchgate = lambda self, target, theta, phi1, phi2: Matchgate(theta, phi1, phi2).apply(self, [target], [])
State.cmatchgate = lambda self, target, theta, phi1, phi2, controls: Matchgate(theta, phi1, phi2).apply(
    self, [target], controls)
State.toffoli = lambda self, c1, c2, target: Toffoli().apply(self, [target], [c1, c2])              # state.rs:2343
State.cnot_ = lambda self, control, target: self.apply_(CNOT(), [target], [control])
State.swap_ = lambda self, q1, q2: self.apply_(SWAP(), [q1, q2], [])
State.toffoli_ = lambda self, c1, c2, target: self.apply_(Toffoli(), [target], [c1, c2])
