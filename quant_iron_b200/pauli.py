"""PauliString and SumOp (src/components/pauli_string.rs) on device states.

A string is lowered to one `qi_pauli_term` record (qubits + Pauli codes + coefficient); the library
turns it into bit masks and runs ONE fused pass for apply, exp or an expectation term, where the
reference makes a clone plus one sweep per factor.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, List

from . import _ffi
from .errors import Error
from .operators import G_X, G_Y, G_Z

_lib = _ffi.lib
_CODE = {G_X: 1, G_Y: 2, G_Z: 3}


def _cmul(a: complex, b: complex) -> complex:
    return complex(a.real * b.real - a.imag * b.imag, a.real * b.imag + a.imag * b.real)


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

    def add_op(self, qubit: int, op):
        if qubit in self._ops:
            raise RuntimeError(f"Duplicate Pauli string operator for qubit: {qubit}")  # panic, pauli_string.rs:66-70
        self._ops[qubit] = op

    def with_op(self, qubit, op):
        self.add_op(qubit, op)
        return self

    def get_targets(self):
        return sorted(self._ops.keys())

    def to_gates(self):  # pauli_string.rs:118-122
        from .circuit import Gate
        return [Gate.Operator(op, [q], []) for q, op in self._ops.items()]

    # ---- C-ABI record ----
    def term(self):
        """(qi_pauli_term, keepalive)"""
        qs = list(self._ops.keys())
        t = _ffi.QiPauliTerm()
        qa = _ffi.u32_array(qs)
        pa = (C.c_uint8 * max(1, len(qs)))(*[_CODE[self._ops[q].KIND] for q in qs])
        t.num_ops = len(qs)
        t.qubits = qa
        t.paulis = pa
        t.coefficient[0], t.coefficient[1] = self._coefficient.real, self._coefficient.imag
        return t, (qa, pa)

    # ---- application (each returns a new State; trailing underscore = in place) ----
    def apply_(self, state, with_coefficient: bool = True):
        t, keep = self.term()
        _ffi.check(_lib.qi_apply_pauli_string(state._h, C.byref(t), 1 if with_coefficient else 0))
        return state

    def apply(self, state):  # pauli_string.rs:139-151
        return self.apply_(state.clone(), True)

    def apply_operators(self, state):  # pauli_string.rs:172-184
        return self.apply_(state.clone(), False)

    def apply_normalised(self, state):  # pauli_string.rs:165-168
        return self.apply_(state.clone(), False).normalise_()

    def apply_exp_factor_(self, state, factor: complex):
        t, keep = self.term()
        f = complex(factor)
        _ffi.check(_lib.qi_apply_pauli_exp(state._h, C.byref(t), _ffi.dbl_array([f.real, f.imag])))
        return state

    def apply_exp(self, state):  # pauli_string.rs:198-223
        return self.apply_exp_factor_(state.clone(), complex(1.0, 0.0))

    def apply_exp_factor(self, state, factor: complex):  # pauli_string.rs:237-262
        return self.apply_exp_factor_(state.clone(), factor)

    def apply_exp_neg_i_dt(self, state, dt: float):  # pauli_string.rs:281-287
        if self._coefficient.imag != 0.0:
            raise Error("InvalidPauliStringCoefficient", self._coefficient)
        return self.apply_exp_factor(state, complex(0.0, -dt))

    def hermitian_conjugate(self):
        return PauliString.with_ops(self._coefficient.conjugate(), self._ops)

    def __mul__(self, rhs):
        return PauliString.with_ops(_cmul(self._coefficient, complex(rhs)), self._ops)

    __rmul__ = __mul__

    def __add__(self, other):
        return SumOp([self, other])

    def __repr__(self):
        ops = " ".join(f"{repr(op)[-1]}[{q}]" for q, op in sorted(self._ops.items()))
        return f"{self._coefficient} * {ops}"


def apply_exp_sequence_(state, strings, factors):
    """In place: state <- exp(factors[k] * c_k * P_k) state for k = 0, 1, ... — the same result as calling
    apply_exp_factor (pauli_string.rs:237-262) once per string, but consecutive strings share fused
    register-window passes on the device (qi_apply_pauli_exp_sequence)."""
    strings = list(strings)
    factors = [complex(f) for f in factors]
    if len(strings) != len(factors):
        raise Error("MismatchedNumberOfParameters", len(strings), len(factors))
    if not strings:
        return state
    arr = (_ffi.QiPauliTerm * len(strings))()
    keep = []
    for i, ps in enumerate(strings):
        rec, k = ps.term()
        arr[i] = rec
        keep.append(k)
    flat = []
    for f in factors:
        flat += [f.real, f.imag]
    _ffi.check(_lib.qi_apply_pauli_exp_sequence(state._h, arr, len(strings), _ffi.dbl_array(flat)))
    return state


class SumOp:
    """pauli_string.rs:398-507."""

    def __init__(self, terms: Iterable[PauliString]):
        self.terms: List[PauliString] = list(terms)

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

    def term_array(self):
        n = len(self.terms)
        arr = (_ffi.QiPauliTerm * max(1, n))()
        keep = []
        for i, t in enumerate(self.terms):
            rec, k = t.term()
            arr[i] = rec
            keep.append(k)
        return arr, n, keep

    def apply(self, state):  # pauli_string.rs:453-466
        from .state import State
        arr, n, keep = self.term_array()
        h = C.c_void_p()
        _ffi.check(_lib.qi_apply_pauli_sum(state._h, arr, n, C.byref(h)))
        return State(_handle=h)

    def expectation_value(self, state) -> complex:  # pauli_string.rs:485-507
        arr, n, keep = self.term_array()
        out = (C.c_double * 2)()
        _ffi.check(_lib.qi_expect_pauli_sum(state._h, arr, n, out))
        return complex(out[0], out[1])

    def __mul__(self, rhs):
        return SumOp([t * rhs for t in self.terms])

    def __add__(self, other):
        if isinstance(other, PauliString):
            return SumOp(self.terms + [other])
        return SumOp(self.terms + other.terms)
