"""Parametric gates (src/components/parametric/{parameter,parametric_gate}.rs).

A `Parameter` is a shared, mutable cell of N floats; a `ParametricGate` holds one and resolves to concrete
`Gate::Operator` gates every time it is applied (gate.rs:107-114), so changing the parameter after a circuit
was built changes what the circuit does.  Host-only: the device sees the resolved operator gates, which join
the fused runs like any other gate.
"""
from __future__ import annotations

import threading
from typing import List, Sequence


class Parameter:
    """parameter.rs:13-74.  `clone()` shares the cell (Rust's `Arc` clone); `deep_clone()` copies it."""

    def __init__(self, initial_values: Sequence[float], _cell=None):
        self._cell = _cell if _cell is not None else [[float(v) for v in initial_values], threading.Lock()]

    @staticmethod
    def new(initial_values: Sequence[float]) -> "Parameter":
        return Parameter(initial_values)

    def clone(self) -> "Parameter":
        return Parameter((), _cell=self._cell)

    def deep_clone(self) -> "Parameter":
        return Parameter(self.get())

    def get(self) -> List[float]:
        with self._cell[1]:
            return list(self._cell[0])

    def set(self, new_values: Sequence[float]) -> None:
        new_values = [float(v) for v in new_values]
        with self._cell[1]:
            if len(new_values) != len(self._cell[0]):
                raise ValueError(f"Parameter<{len(self._cell[0])}>::set got {len(new_values)} values")   # a type error in Rust
            self._cell[0] = new_values

    def __len__(self):
        return len(self._cell[0])

    def __repr__(self):
        return f"Parameter({self.get()})"


class ParametricGate:
    """`trait ParametricGate` (parametric_gate.rs:6-24)."""
    ARITY = 1

    def __init__(self, parameter: Parameter):
        if len(parameter) != self.ARITY:
            raise ValueError(f"{type(self).__name__} takes a Parameter of {self.ARITY} values")
        self.parameter = parameter

    def to_concrete_gates(self, target_indices, control_indices):
        raise NotImplementedError

    def box_clone(self):
        return type(self)(self.parameter.clone())

    def __repr__(self):
        return f"{type(self).__name__} {{ parameter: {self.parameter!r} }}"


def _gate():
    from .circuit import Gate
    return Gate


class ParametricRyPhase(ParametricGate):      # parametric_gate.rs:35-52
    ARITY = 2

    def to_concrete_gates(self, target_indices, control_indices):
        th, ph = self.parameter.get()
        return _gate().ry_phase_controlled_gates(list(target_indices), list(control_indices), th, ph)


class ParametricRyPhaseDag(ParametricGate):   # parametric_gate.rs:63-80
    ARITY = 2

    def to_concrete_gates(self, target_indices, control_indices):
        th, ph = self.parameter.get()
        return _gate().ry_phase_dag_controlled_gates(list(target_indices), list(control_indices), th, ph)


class ParametricMatchgate(ParametricGate):    # parametric_gate.rs:92-110
    ARITY = 3

    def to_concrete_gates(self, target_indices, control_indices):
        th, p1, p2 = self.parameter.get()
        return [_gate().controlled_matchgate(target_indices[0], list(control_indices), th, p1, p2)]


class ParametricRx(ParametricGate):           # parametric_gate.rs:120-136
    def to_concrete_gates(self, target_indices, control_indices):
        return _gate().rx_controlled_gates(list(target_indices), list(control_indices), self.parameter.get()[0])


class ParametricRy(ParametricGate):           # parametric_gate.rs:146-162
    def to_concrete_gates(self, target_indices, control_indices):
        return _gate().ry_controlled_gates(list(target_indices), list(control_indices), self.parameter.get()[0])


class ParametricRz(ParametricGate):           # parametric_gate.rs:172-188
    def to_concrete_gates(self, target_indices, control_indices):
        return _gate().rz_controlled_gates(list(target_indices), list(control_indices), self.parameter.get()[0])


class ParametricP(ParametricGate):            # parametric_gate.rs:198-211
    def to_concrete_gates(self, target_indices, control_indices):
        return _gate().p_controlled_gates(list(target_indices), list(control_indices), self.parameter.get()[0])
