"""Trotterised time evolution (src/algorithms/time_evolution.rs) on device states."""
from __future__ import annotations

from . import _ffi
from .errors import Error
from .pauli import SumOp

_lib = _ffi.lib


class TrotterOrder:
    First = "First"
    Second = "Second"


def _evolve(hamiltonian: SumOp, state, dt: float, steps: int, order: int):
    out = state.clone()
    arr, n, keep = hamiltonian.term_array()
    _ffi.check(_lib.qi_trotter_evolve(out._h, arr, n, float(dt), int(steps), order))
    return out


def first_order_trotter_step(hamiltonian: SumOp, initial_state, dt: float):   # time_evolution.rs:45-66
    return _evolve(hamiltonian, initial_state, dt, 1, 1)


def second_order_trotter_step(hamiltonian: SumOp, initial_state, dt: float):  # time_evolution.rs:89-115
    return _evolve(hamiltonian, initial_state, dt, 1, 2)


def trotter_evolve_state(hamiltonian: SumOp, initial_state, dt: float, num_steps: int, order):  # 140-167
    if hamiltonian.num_terms() == 0:
        raise Error("InvalidNumberOfQubits", 0)
    return _evolve(hamiltonian, initial_state, dt, num_steps, 1 if order == TrotterOrder.First else 2)


def trotter_evolve_state_(hamiltonian: SumOp, state, dt: float, num_steps: int, order):
    """In-place variant (no clone of the state)."""
    arr, n, keep = hamiltonian.term_array()
    _ffi.check(_lib.qi_trotter_evolve(state._h, arr, n, float(dt), int(num_steps),
                                      1 if order == TrotterOrder.First else 2))
    return state
