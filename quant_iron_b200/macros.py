"""`circuit!` (src/macros.rs:99-380) as a Python DSL.

The Rust macro is a token muncher that turns `circuit!{ qubits: n, h(0), cx([0, 1], 2), ... }` into CircuitBuilder
calls followed by `build_final()`.  Host-only sugar; here the operations are values:

    from quant_iron_b200.macros import circuit, ops as q
    c = circuit(3, q.h(0), q.cx([0, 1], 2), q.rx([1, 2], 0.5), q.measurez([0, 1]))

Every rule expands exactly as the macro's does (cited per group below), including its overloading on a single index vs a
bracketed list and its argument-order quirks: `cnot(a, b)` calls `cnot_gate(a, b)` whose parameters are
(target, control) (macros.rs:141, circuit.rs:1071); `toffoli(a, b, c)` calls `toffoli_gate(a, b, c)` whose parameters are
(control1, control2, target) (macros.rs:227, circuit.rs:1118-1123); `cmatchgate(t, c, ...)` passes the controls second
(macros.rs:231-232, circuit.rs:1194-1200).  Errors surface from `build_final()` (circuit.rs:352) as in the macro; a
non-unitary matrix raises at once (the macro unwraps, macros.rs:216-222).
"""
from __future__ import annotations

from typing import Any, Callable, Sequence

from .measurement import MeasurementBasis

_SINGLE = ("h", "x", "y", "z", "s", "t", "id", "sdag", "tdag")                  # macros.rs:117-137
_CONTROLLED = ("ch", "cx", "cy", "cz", "cs", "csdag", "ct", "ctdag")            # macros.rs:144-175
_ANGLE = ("rx", "ry", "rz", "p")                                                # macros.rs:177-186
_ANGLE2 = ("ry_phase", "ry_phase_dag")                                          # macros.rs:181-188
_CONTROLLED_ANGLE = ("crx", "cry", "crz", "cp")                                 # macros.rs:190-205
_CONTROLLED_ANGLE2 = ("cry_phase", "cry_phase_dag")                             # macros.rs:206-213
_NAMES = set(_SINGLE + _CONTROLLED + _ANGLE + _ANGLE2 + _CONTROLLED_ANGLE + _CONTROLLED_ANGLE2 + (
    "cnot", "swap", "unitary", "cunitary", "toffoli", "cswap", "matchgate", "cmatchgate", "pauli_string",
    "pauli_time_evolution", "measurex", "measurey", "measurez", "measure_custom"))


class Op:
    """One `name(args...)` item of the macro body."""

    def __init__(self, name: str, args: tuple):
        self.name, self.args = name, args

    def __repr__(self):
        return f"{self.name}({', '.join(map(repr, self.args))})"


class _Ops:
    def __getattr__(self, name: str) -> Callable[..., Op]:
        if name not in _NAMES:
            raise AttributeError(f"circuit!: no rule for `{name}`")          # the macro fails to compile
        return lambda *args: Op(name, args)


ops = _Ops()


def _is_list(a: Any) -> bool:
    return isinstance(a, (list, tuple))


def _vec(a: Any) -> list:
    return list(a) if _is_list(a) else [a]


def _arity(op: Op, n: int):
    if len(op.args) != n:
        raise TypeError(f"circuit!: no rule matches `{op!r}`")


def expand(builder, operations: Sequence[Op]):
    """circuit_internal! (macros.rs:111-380): one builder call per operation, in order."""
    for op in operations:
        n, a = op.name, op.args
        if n in _SINGLE:
            _arity(op, 1)
            getattr(builder, f"{n}_gates" if _is_list(a[0]) else f"{n}_gate")(list(a[0]) if _is_list(a[0]) else a[0])
        elif n in ("cnot", "swap"):                                            # macros.rs:141-142
            _arity(op, 2)
            getattr(builder, f"{n}_gate")(a[0], a[1])
        elif n in _CONTROLLED:
            _arity(op, 2)
            getattr(builder, f"{n}_gates")(_vec(a[0]), _vec(a[1]))
        elif n in _ANGLE:
            _arity(op, 2)
            getattr(builder, f"{n}_gates" if _is_list(a[0]) else f"{n}_gate")(list(a[0]) if _is_list(a[0]) else a[0], a[1])
        elif n in _ANGLE2:
            _arity(op, 3)
            getattr(builder, f"{n}_gates" if _is_list(a[0]) else f"{n}_gate")(list(a[0]) if _is_list(a[0]) else a[0], a[1], a[2])
        elif n in _CONTROLLED_ANGLE:
            _arity(op, 3)
            getattr(builder, f"{n}_gates")(_vec(a[0]), _vec(a[1]), a[2])
        elif n in _CONTROLLED_ANGLE2:
            _arity(op, 4)
            getattr(builder, f"{n}_gates")(_vec(a[0]), _vec(a[1]), a[2], a[3])
        elif n == "unitary":                                                   # macros.rs:215-216
            _arity(op, 2)
            (builder.unitary_gates if _is_list(a[0]) else builder.unitary_gate)(list(a[0]) if _is_list(a[0]) else a[0], a[1])
        elif n == "cunitary":                                                  # macros.rs:217-220
            _arity(op, 3)
            builder.cunitary_gates(_vec(a[0]), _vec(a[1]), a[2])
        elif n == "toffoli":                                                   # macros.rs:227
            _arity(op, 3)
            builder.toffoli_gate(a[0], a[1], a[2])
        elif n == "cswap":                                                     # macros.rs:228-229
            _arity(op, 3)
            builder.cswap_gate(a[0], a[1], _vec(a[2]))
        elif n == "matchgate":                                                 # macros.rs:230
            _arity(op, 4)
            builder.matchgate(a[0], a[1], a[2], a[3])
        elif n == "cmatchgate":                                                # macros.rs:231-232
            _arity(op, 5)
            builder.cmatchgate(a[0], _vec(a[1]), a[2], a[3], a[4])
        elif n == "pauli_string":                                              # macros.rs:233
            _arity(op, 1)
            builder.pauli_string_gate(a[0])
        elif n == "pauli_time_evolution":                                      # macros.rs:234
            _arity(op, 2)
            builder.pauli_time_evolution_gate(a[0], a[1])
        elif n in ("measurex", "measurey", "measurez"):                        # macros.rs:236-241
            _arity(op, 1)
            builder.measure_gate(_basis_for(builder, {"measurex": "X", "measurey": "Y", "measurez": "Computational"}[n]), _vec(a[0]))
        elif n == "measure_custom":                                            # macros.rs:242-243
            _arity(op, 2)
            builder.measure_gate(_custom_basis_for(builder, a[1]), _vec(a[0]))
        else:
            raise TypeError(f"circuit!: no rule for `{n}`")
    return builder


def _basis_for(builder, name: str):
    """The MeasurementBasis of the implementation the builder belongs to (this package, or the oracle in tests)."""
    mod = __import__(type(builder).__module__, fromlist=["MeasurementBasis"])
    return getattr(getattr(mod, "MeasurementBasis", MeasurementBasis), name)


def _custom_basis_for(builder, matrix):
    mod = __import__(type(builder).__module__, fromlist=["MeasurementBasis"])
    return getattr(mod, "MeasurementBasis", MeasurementBasis).Custom(matrix)


def circuit(qubits: int, *operations: Op, builder_cls=None):
    """circuit!{ qubits: n, ops... } (macros.rs:99-108): CircuitBuilder::new(n), the expanded calls, build_final()."""
    if builder_cls is None:
        from .circuit import CircuitBuilder as builder_cls
    return expand(builder_cls(qubits), operations).build_final()
