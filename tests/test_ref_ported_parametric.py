"""Port of src/tests/parametric_tests.rs: Parameter cells, parametric gates resolving to concrete gates, circuits
whose behaviour follows the parameter after they were built.  The reference compares `format!("{:?}")` of the
concrete circuits; here a gate is described by (operator type, numeric parameters, targets, controls) and, because
this is an engine test, the circuits are also EXECUTED and compared amplitude for amplitude.
Runs against the CPU oracle everywhere and against the GPU engine on a B200."""
import pytest

from conftest import assert_amps, vec


def describe(circuit):
    return [(type(g.op).__name__, tuple(g.op.params()), tuple(g.targets), tuple(g.controls)) for g in circuit.gates]


def same(qi, parametric_circuit, concrete_circuit, n):
    assert describe(parametric_circuit.to_concrete_circuit()) == describe(concrete_circuit)
    st = qi.State.new_plus(n).t(0)
    assert_amps(parametric_circuit.execute(st), vec(concrete_circuit.execute(st)), msg="parametric vs concrete execution")


def test_parameter_cell(qi):
    """parametric_tests.rs:5-39."""
    p = qi.Parameter.new([0.5, 1.0])
    assert p.get() == [0.5, 1.0]
    p.set([0.6, 1.1])
    assert p.get() == [0.6, 1.1]
    p = qi.Parameter.new([0.5, 1.0])
    shared, deep = p.clone(), p.deep_clone()
    assert shared.get() == [0.5, 1.0] and deep.get() == [0.5, 1.0]
    p.set([0.6, 1.1])
    assert shared.get() == [0.6, 1.1] and deep.get() == [0.5, 1.0]


def test_parametric_ry_phase(qi):
    """parametric_tests.rs:41-77."""
    p = qi.Parameter.new([0.5, 1.0])
    th, ph = p.get()
    same(qi, qi.CircuitBuilder(1).parametric_ry_phase_gate(0, p.clone()).build_final(),
         qi.CircuitBuilder(1).ry_phase_gate(0, th, ph).build_final(), 1)
    same(qi, qi.CircuitBuilder(3).parametric_cry_phase_gates([0, 1], [2], [p.clone(), p.clone()]).build_final(),
         qi.CircuitBuilder(3).cry_phase_gates([0, 1], [2], th, ph).build_final(), 3)
    same(qi, qi.CircuitBuilder(2).parametric_ry_phase_gates([0, 1], [p.clone(), p.clone()]).build_final(),
         qi.CircuitBuilder(2).ry_phase_gates([0, 1], th, ph).build_final(), 2)


def test_parametric_ry_phase_dag(qi):
    """parametric_tests.rs:388-444."""
    p = qi.Parameter.new([0.5, 1.0])
    th, ph = p.get()
    same(qi, qi.CircuitBuilder(1).parametric_ry_phase_dag_gate(0, p.clone()).build_final(),
         qi.CircuitBuilder(1).ry_phase_dag_gate(0, th, ph).build_final(), 1)
    same(qi, qi.CircuitBuilder(2).parametric_ry_phase_dag_gates([0, 1], [p.clone(), p.clone()]).build_final(),
         qi.CircuitBuilder(2).ry_phase_dag_gates([0, 1], th, ph).build_final(), 2)
    same(qi, qi.CircuitBuilder(3).parametric_cry_phase_dag_gates([0, 1], [2], [p.clone(), p.clone()]).build_final(),
         qi.CircuitBuilder(3).cry_phase_dag_gates([0, 1], [2], th, ph).build_final(), 3)


def test_parametric_matchgate(qi):
    """parametric_tests.rs:79-104."""
    p = qi.Parameter.new([0.5, 1.0, 1.5])
    a, b, c = p.get()
    same(qi, qi.CircuitBuilder(2).parametric_matchgate(0, p.clone()).build_final(),
         qi.CircuitBuilder(2).matchgate(0, a, b, c).build_final(), 2)
    same(qi, qi.CircuitBuilder(3).parametric_cmatchgate(0, [2], p.clone()).build_final(),
         qi.CircuitBuilder(3).cmatchgate(0, [2], a, b, c).build_final(), 3)


def test_parametric_change_parameter_value_after_build(qi):
    """parametric_tests.rs:106-135: the built circuit follows later `set` calls."""
    ryp, mc = qi.Parameter.new([0.5, 1.0]), qi.Parameter.new([0.5, 1.0, 1.5])
    circuit = (qi.CircuitBuilder(3).parametric_cry_phase_gates([0, 1], [2], [ryp.clone(), ryp.clone()])
               .parametric_matchgate(0, mc.clone()).build_final())

    def concrete():
        th, ph = ryp.get()
        a, b, c = mc.get()
        return qi.CircuitBuilder(3).cry_phase_gates([0, 1], [2], th, ph).matchgate(0, a, b, c).build_final()
    same(qi, circuit, concrete(), 3)
    st = qi.State.new_plus(3).t(0)
    before = vec(circuit.execute(st)).copy()
    ryp.set([0.6, 1.1])
    mc.set([0.6, 1.1, 1.6])
    same(qi, circuit, concrete(), 3)
    assert abs(vec(circuit.execute(st)) - before).max() > 1e-3      # the same circuit object now does something else


def test_parametric_mismatched_parameters(qi):
    """parametric_tests.rs:137-244: every multi-target adder checks len(targets) == len(parameters)."""
    p2, p1 = qi.Parameter.new([0.5, 1.0]), qi.Parameter.new([0.5])
    cases = [
        lambda b: b.parametric_cry_phase_gates([0, 1], [2], [p2.clone()]),
        lambda b: b.parametric_ry_phase_gates([0, 1], [p2.clone()]),
        lambda b: b.parametric_crx_gates([0, 1], [2], [p1.clone()]),
        lambda b: b.parametric_cry_gates([0, 1], [2], [p1.clone()]),
        lambda b: b.parametric_crz_gates([0, 1], [2], [p1.clone()]),
        lambda b: b.parametric_cp_gates([0, 1], [2], [p1.clone()]),
        lambda b: b.parametric_cry_phase_dag_gates([0, 1], [2], [p2.clone()]),
        lambda b: b.parametric_ry_phase_dag_gates([0, 1], [p2.clone()]),
        lambda b: b.parametric_rx_gates([0, 1], [p1.clone()]),
        lambda b: b.parametric_p_gates([0, 1], [p1.clone()]),
    ]
    for add in cases:
        with pytest.raises(qi.Error) as e:
            add(qi.CircuitBuilder(3))
        assert e.value.variant == "MismatchedNumberOfParameters" and e.value.payload == (2, 1)
        assert e.value.to_string() == "Mismatched number of parameters: expected 2, got 1"


@pytest.mark.parametrize("name", ["rx", "ry", "rz", "p"])
def test_parametric_rotations_and_phase(qi, name):
    """parametric_tests.rs:246-386: single, multi and controlled forms of Rx / Ry / Rz / P."""
    p = qi.Parameter.new([0.5])
    a = p.get()[0]
    B = qi.CircuitBuilder
    same(qi, getattr(B(1), f"parametric_{name}_gate")(0, p.clone()).build_final(), getattr(B(1), f"{name}_gate")(0, a).build_final(), 1)
    same(qi, getattr(B(2), f"parametric_c{name}_gates")([0], [1], [p.clone()]).build_final(),
         getattr(B(2), f"c{name}_gates")([0], [1], a).build_final(), 2)
    same(qi, getattr(B(2), f"parametric_{name}_gates")([0, 1], [p.clone(), p.clone()]).build_final(),
         getattr(B(2), f"{name}_gates")([0, 1], a).build_final(), 2)
    # Gate::apply on a parametric gate folds the concrete gates over the state (gate.rs:107-114)
    g = qi.Gate.Parametric({"rx": qi.ParametricRx, "ry": qi.ParametricRy, "rz": qi.ParametricRz, "p": qi.ParametricP}[name](p.clone()), [0], [1])
    st = qi.State.new_plus(2)
    assert_amps(g.apply(st), vec(getattr(B(2), f"c{name}_gates")([0], [1], a).build_final().execute(st)))
