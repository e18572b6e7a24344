"""The reference's `circuit!` tests (src/tests/macros_tests.rs, all 13) against the Python form of the macro
(quant_iron_b200/macros.py), expanded through the CircuitBuilder of either implementation: "oracle" everywhere, the
device engine's host mirror under `-m gpu` (building a circuit needs no device, running it does)."""
import math

import numpy as np
import pytest

from quant_iron_b200.macros import circuit, ops as q

PI = math.pi
X_MATRIX = [[0.0, 1.0], [1.0, 0.0]]


def _c(qi, n, *operations):
    return circuit(n, *operations, builder_cls=qi.CircuitBuilder)


def _err(qi, variant, payload, n, *operations):
    with pytest.raises(qi.Error) as e:
        _c(qi, n, *operations)
    assert e.value.variant == variant and tuple(e.value.payload) == payload, e.value


SINGLE = lambda last: [q.h(0), q.x(1), q.y(2), q.z(0), q.s(1), q.t(2), q.id(0), q.sdag(1), q.tdag(last)]
MIXED = lambda last: [q.h(0), q.h([0, 1]), q.x(1), q.x([1, 2]), q.y(2), q.y([0, 2]), q.z(0), q.z([0, 1]), q.s(1), q.s([1, 2]),
                      q.t(2), q.t([0, 2]), q.id(0), q.id([0, 1]), q.sdag(1), q.sdag([1, 2]), q.tdag(2), q.tdag([0, last])]


def test_circuit_macro_single_qubit_gates_success(qi):  # macros_tests.rs:7-25
    c = _c(qi, 3, *SINGLE(2))
    assert c.num_qubits == 3 and len(c.gates) == 9


def test_circuit_macro_single_qubit_gates_failure(qi):  # macros_tests.rs:27-45
    _err(qi, "InvalidQubitIndex", (3, 3), 3, *SINGLE(3))


def test_circuit_macro_mixed_multi_qubit_gates_success(qi):  # macros_tests.rs:47-74
    c = _c(qi, 3, *MIXED(2))
    assert c.num_qubits == 3 and len(c.gates) == 9 * 3


def test_circuit_macro_mixed_multi_qubit_gates_failure(qi):  # macros_tests.rs:76-103
    _err(qi, "InvalidQubitIndex", (3, 3), 3, *MIXED(3))


def test_circuit_macro_two_three_qubit_gates_success(qi):  # macros_tests.rs:105-118
    c = _c(qi, 5, q.cnot(0, 1), q.swap(1, 2), q.cswap(0, 1, 2), q.cswap(0, 1, [2, 3]), q.toffoli(0, 1, 2))
    assert c.num_qubits == 5 and len(c.gates) == 5


def test_circuit_macro_two_three_qubit_gates_failure(qi):  # macros_tests.rs:120-132
    _err(qi, "InvalidQubitIndex", (6, 5), 5, q.cnot(0, 1), q.swap(1, 2), q.cswap(0, 1, 2), q.cswap(0, 1, [2, 6]), q.toffoli(0, 1, 2))


def _four(f, *extra):
    return [f(0, 1, *extra), f([0, 1], 2, *extra), f(0, [1, 2], *extra), f([0, 1], [2, 3], *extra)]


def test_circuit_macro_controlled_gates_success(qi):  # macros_tests.rs:134-211
    body = []
    for f in (q.ch, q.cx, q.cy, q.cz, q.cs, q.ct, q.csdag, q.ctdag):
        body += _four(f)
    body += _four(q.crx, PI / 4) + _four(q.cry, PI / 4) + _four(q.cry_phase, PI / 4, PI / 2)
    body += [q.cmatchgate(0, 1, PI / 4, PI / 2, PI / 3), q.cmatchgate(0, [1, 2], PI / 4, PI / 2, PI / 3)]
    body += _four(q.crz, PI / 4) + _four(q.cp, PI / 4)
    c = _c(qi, 4, *body)
    assert c.num_qubits == 4 and len(c.gates) == 6 * 13 + 2


def test_circuit_macro_controlled_gates_failure(qi):  # macros_tests.rs:213-229
    _err(qi, "InvalidQubitIndex", (6, 4), 4, *(_four(q.ch) + [q.cx(0, 1), q.cx([0, 1], 2), q.cx(0, [1, 2]), q.cx([0, 1], [2, 6])]))


def test_circuit_macro_unitary_cunitary_gates_success(qi):  # macros_tests.rs:231-247
    c = _c(qi, 4, q.unitary(0, X_MATRIX), q.unitary([0, 1], X_MATRIX), *_four(q.cunitary, X_MATRIX))
    assert c.num_qubits == 4 and len(c.gates) == 9


def test_circuit_macro_unitary_cunitary_gates_failure(qi):  # macros_tests.rs:249-265
    _err(qi, "InvalidQubitIndex", (6, 4), 4, q.unitary(0, X_MATRIX), q.unitary([0, 1], X_MATRIX), q.cunitary(0, 1, X_MATRIX),
         q.cunitary([0, 1], 2, X_MATRIX), q.cunitary(0, [1, 2], X_MATRIX), q.cunitary([0, 1], [2, 6], X_MATRIX))


def test_circuit_macro_measurement_success(qi):  # macros_tests.rs:267-284
    c = _c(qi, 3, q.measurex(0), q.measurex([1, 2]), q.measurey(0), q.measurey([1, 2]), q.measurez(0), q.measurez([1, 2]),
           q.measure_custom(0, X_MATRIX), q.measure_custom([1, 2], X_MATRIX))
    assert c.num_qubits == 3 and len(c.gates) == 8


def test_circuit_macro_angle_gates_success(qi):  # macros_tests.rs:286-305
    c = _c(qi, 3, q.rx(0, PI / 2), q.ry(1, PI / 2), q.rz(2, PI / 2), q.p(0, PI / 2), q.ry_phase(1, PI / 2, PI / 4),
           q.matchgate(0, PI / 2, PI / 3, PI / 4), q.rx([0, 1], PI / 2), q.ry([1, 2], PI / 2), q.rz([0, 2], PI / 2),
           q.ry_phase([0, 1], PI / 2, PI / 4), q.p([0, 1, 2], PI / 2))
    assert c.num_qubits == 3 and len(c.gates) == 4 + 2 + 2 + 2 + 3 + 4


def test_circuit_macro_match_gates_success(qi):  # macros_tests.rs:307-323
    c = _c(qi, 4, q.matchgate(0, PI / 2, PI / 3, PI / 4), q.cmatchgate(0, 2, PI / 2, PI / 3, PI / 4),
           q.cmatchgate(0, [2, 3], PI / 2, PI / 3, PI / 4))
    assert c.num_qubits == 4 and len(c.gates) == 3


def test_macro_argument_order_quirks_and_execution(qi):
    """cnot(a, b) has a = TARGET; toffoli(a, b, c) has c = TARGET; the expanded circuit runs."""
    c = _c(qi, 3, q.x(1), q.cnot(0, 1))                       # control 1 is set -> target 0 flips: |010> -> |011>
    out = c.execute(qi.State.new_zero(3))
    assert abs(out.amplitude(3) - 1.0) < 1e-12
    c = _c(qi, 3, q.x([0, 1]), q.toffoli(0, 1, 2))            # controls 0, 1 set -> target 2 flips: |011> -> |111>
    assert abs(c.execute(qi.State.new_zero(3)).amplitude(7) - 1.0) < 1e-12
    with pytest.raises(AttributeError):
        q.hadamard(0)                                          # no such rule: the Rust macro would not compile
    with pytest.raises(TypeError):
        _c(qi, 2, q.rx(0))                                     # wrong arity: no rule matches
