"""Port of the remaining reference tests that the larger ported files did not restate: Gate constructors and
accessors (gate_tests.rs:11-60), CircuitBuilder life cycle (circuit_tests.rs:159-255), PauliString::get_targets /
to_gates (pauli_string_tests.rs:366-400), measure_n over all qubits and its errors (measurement_tests.rs:152-207).
Not ported: `circuit!` macro tests (macros_tests.rs, Rust macro sugar) and to_qasm (circuit_tests.rs:134-157,
string emission) -- both out of scope.  Runs on the CPU oracle everywhere and on the GPU engine."""
import numpy as np
import pytest

from conftest import vec


def _raises(qi, variant, payload, fn):
    with pytest.raises(qi.Error) as e:
        fn()
    assert e.value.variant == variant and tuple(e.value.payload) == tuple(payload), e.value


def test_gate_constructors_and_targets(qi):
    """gate_tests.rs:11-60."""
    G, P = qi.Gate, qi.Pauli
    g = G.Operator(qi.Hadamard(), [0, 1], [])
    assert g.get_target_qubits() == [0, 1] and g.get_control_qubits() == []
    m = G.Measurement(qi.MeasurementBasis.Computational, [0, 2])
    assert m.get_target_qubits() == [0, 2] and m.get_control_qubits() is None
    pg = G.Parametric(qi.ParametricP(qi.Parameter.new([0.5])), [0], [])
    assert pg.get_target_qubits() == [0] and pg.get_control_qubits() == []
    ps = qi.PauliString.new(1.0).with_op(1, P.X).with_op(2, P.Y)
    assert sorted(G.PauliString(ps).get_target_qubits()) == [1, 2]
    ev = G.PauliTimeEvolution(qi.PauliString.new(1.0).with_op(0, P.Z).with_op(1, P.X), 0.01)
    assert sorted(ev.get_target_qubits()) == [0, 1] and ev.get_control_qubits() is None


def test_circuit_builder_life_cycle(qi):
    """circuit_tests.rs:159-255."""
    B, G = qi.CircuitBuilder, qi.Gate
    b = B.new(3)
    assert b.num_qubits == 3 and b.gates == []
    b = B.new(2)
    b.add_gate(G.h_gate(1))
    assert len(b.gates) == 1
    b = B.new(2)
    b.add_gates([G.h_gate(1), G.cnot_gate(0, 1)])
    assert len(b.gates) == 2
    # build keeps the gates in the builder, build_final drains it
    b = B.new(2)
    c = b.h_gate(0).cnot_gate(0, 1).build()
    assert c.num_qubits == 2 and len(c.gates) == 2 and len(b.gates) == 2
    b = B.new(2)
    c = b.h_gate(0).cnot_gate(0, 1).build_final()
    assert c.num_qubits == 2 and len(c.gates) == 2 and b.gates == []
    # an out-of-range qubit is reported when the circuit is built (circuit.rs:35-52)
    _raises(qi, "InvalidQubitIndex", (3, 2), lambda: B.new(2).h_gate(0).cnot_gate(0, 3).build())
    _raises(qi, "InvalidQubitIndex", (3, 2), lambda: B.new(2).h_gate(0).cnot_gate(0, 3).build_final())
    # subroutines
    b = B.new(2)
    sub = b.h_gate(0).cnot_gate(0, 1).build_subroutine()
    assert sub.num_qubits == 2 and len(sub.gates) == 2
    b.add_subroutine(sub)
    assert len(b.gates) == 2
    bell = b.build().execute(qi.State.new_zero(2))
    assert bell == qi.State.new_zero(2).h(0).cnot(1, 0) or bell == qi.State.new_zero(2).h(0).cnot(0, 1)


def test_pauli_string_targets_and_gates(qi):
    """pauli_string_tests.rs:366-400."""
    P = qi.Pauli
    ps = qi.PauliString.new(1.0)
    ps.add_op(2, P.X)
    ps.add_op(0, P.Y)
    ps.add_op(1, P.Z)
    assert sorted(ps.get_targets()) == [0, 1, 2]
    gates = ps.to_gates()
    assert len(gates) == 3
    got = sorted((g.get_target_qubits()[0], repr(g.op)) for g in gates)
    exp = sorted((g.get_target_qubits()[0], repr(g.op)) for g in (qi.Gate.x_gate(2), qi.Gate.y_gate(0), qi.Gate.z_gate(1)))
    assert got == exp == [(0, "Pauli.Y"), (1, "Pauli.Z"), (2, "Pauli.X")]
    # and they act like the string itself (up to the coefficient, which is 1 here)
    st = qi.State.new_plus(3).t(1)
    out = st
    for g in gates:
        out = g.apply(out)
    assert out == ps.apply(st)


def test_measure_n_all_qubits_and_errors(qi):
    """measurement_tests.rs:152-207: empty qubit list = every qubit; each result collapses onto its outcome."""
    st = qi.State.new([0.5 + 0j] * 4)
    results = st.measure_n(qi.MeasurementBasis.Computational, [], 5, seed=11)
    assert len(results) == 5 and results[0].get_basis() == qi.MeasurementBasis.Computational
    for r in results:
        o0, o1 = r.get_outcomes()[0], r.get_outcomes()[1]
        exp = np.zeros(4, dtype=np.complex128)
        exp[(o1 << 1) | o0] = 1.0
        assert np.allclose(vec(r.get_new_state()), exp, atol=1e-15) and r.get_new_state().num_qubits == 2
    M = qi.MeasurementBasis.Computational
    _raises(qi, "InvalidQubitIndex", (3, 2), lambda: st.measure_n(M, [3], 5))
    _raises(qi, "InvalidNumberOfQubits", (2,), lambda: st.measure_n(M, [0, 1, 2], 5))
    _raises(qi, "InvalidNumberOfMeasurements", (0,), lambda: st.measure_n(M, [0], 0))
