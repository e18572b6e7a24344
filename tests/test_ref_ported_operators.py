"""Port of the reference's per-operator known-answer tests (src/tests/operator_tests.rs).

Each test cites the reference test it restates.  The facts are the reference's; expected states
are built from the textbook 2x2 matrices with numpy (not with the implementation under test).
They run against the CPU oracle everywhere (pinning it) and against the GPU engine on a B200.
Equality is checked twice: with the reference's own `State == State` (f32::EPSILON per component,
state.rs:2360-2367) and with this build's 1e-12 amplitude bar.
The reference's `#[cfg(feature="gpu")]` 15-qubit cases are ported too (N_BIG).
"""
import cmath
import math

import numpy as np
import pytest

from conftest import assert_amps, basis, vec

PI = math.pi
S2 = 1.0 / math.sqrt(2.0)
N_PAR = 11     # reference's rayon-path size (operator.rs:18)
N_BIG = 15     # reference's OpenCL-path size (operator.rs:21)

ZERO = np.array([1, 0], dtype=complex)
ONE = np.array([0, 1], dtype=complex)
PLUS = np.array([S2, S2], dtype=complex)
MINUS = np.array([S2, -S2], dtype=complex)


def M_H():
    return np.array([[S2, S2], [S2, -S2]], dtype=complex)


def M_X():
    return np.array([[0, 1], [1, 0]], dtype=complex)


def M_Y():
    return np.array([[0, -1j], [1j, 0]], dtype=complex)


def M_Z():
    return np.array([[1, 0], [0, -1]], dtype=complex)


def M_P(t):
    return np.array([[1, 0], [0, cmath.exp(1j * t)]], dtype=complex)


def M_RX(t):
    c, s = math.cos(t / 2), math.sin(t / 2)
    return np.array([[c, -1j * s], [-1j * s, c]], dtype=complex)


def M_RY(t):
    c, s = math.cos(t / 2), math.sin(t / 2)
    return np.array([[c, -s], [s, c]], dtype=complex)


def M_RZ(t):
    return np.array([[cmath.exp(-1j * t / 2), 0], [0, cmath.exp(1j * t / 2)]], dtype=complex)


def kron_all(vs):
    """vs[q] is the 1-qubit state of qubit q (qubit 0 = least significant bit)."""
    out = np.array([1.0 + 0j])
    for v in vs:
        out = np.kron(v, out)
    return out


THETA = PI / 2.5

# name -> (matrix, call-args, reference test lines)
SINGLE = {
    "h": (M_H(), (), "operator_tests.rs:16-121"),
    "x": (M_X(), (), "operator_tests.rs:123-237"),
    "y": (M_Y(), (), "operator_tests.rs:239-337"),
    "z": (M_Z(), (), "operator_tests.rs:339-448"),
    "i": (np.eye(2, dtype=complex), (), "operator_tests.rs:486-540"),
    "s": (M_P(PI / 2), (), "operator_tests.rs:542-718"),
    "t": (M_P(PI / 4), (), "operator_tests.rs:720-856"),
    "s_dag": (M_P(-PI / 2), (), "operator_tests.rs:858-984"),
    "t_dag": (M_P(-PI / 4), (), "operator_tests.rs:986-1106"),
    "p": (M_P(THETA), (THETA,), "operator_tests.rs:1108-1248"),
    "rx": (M_RX(THETA), (THETA,), "operator_tests.rs:1250-1426"),
    "ry": (M_RY(THETA), (THETA,), "operator_tests.rs:1428-1590"),
    "rz": (M_RZ(THETA), (THETA,), "operator_tests.rs:1592-1750"),
}


def _state(qi, v):
    n = int(math.log2(len(v)))
    return qi.State(np.asarray(v, dtype=np.complex128), n)


@pytest.mark.parametrize("name", sorted(SINGLE))
def test_single_qubit_known_answers(qi, name):
    """<gate>(|0>), (|1>), (|+>), (|->) and <gate>_multi(|00>): first block of every
    test_operator_*_success (e.g. operator_tests.rs:22-39, 129-147, 1276-1288)."""
    m, args, _cite = SINGLE[name]
    for v in (ZERO, ONE, PLUS, MINUS):
        out = getattr(_state(qi, v), name)(0, *args)
        assert_amps(out, m @ v, msg=f"{name} on {v}")
        assert out == _state(qi, m @ v)
    out = getattr(qi.State.new_zero(2), f"{name}_multi")([0, 1], *args)
    assert_amps(out, kron_all([m @ ZERO, m @ ZERO]), msg=f"{name}_multi |00>")


@pytest.mark.parametrize("name", sorted(SINGLE))
def test_single_qubit_controlled_two_qubit(qi, name):
    """c<gate>(control=0,target=1) on |psi>|1> applies the gate, on control=0 leaves the state
    (e.g. operator_tests.rs:41-49, 154-165, 1290-1299)."""
    m, args, _ = SINGLE[name]
    for tv in (ZERO, ONE, PLUS):
        st = _state(qi, kron_all([ONE, tv]))            # qubit0 = |1> (control), qubit1 = tv
        out = getattr(st, f"c{name}_multi")([1], [0], *args)
        assert_amps(out, kron_all([ONE, m @ tv]), msg=f"c{name} control=1")
        st0 = _state(qi, kron_all([ZERO, tv]))
        out0 = getattr(st0, f"c{name}_multi")([1], [0], *args)
        assert_amps(out0, vec(st0), msg=f"c{name} control=0")


@pytest.mark.parametrize("n", [N_PAR, N_BIG])
@pytest.mark.parametrize("name", sorted(SINGLE))
def test_single_qubit_parallel_path_sizes(qi, name, n):
    """11-qubit (rayon path) and 15-qubit (OpenCL path) cases: <gate>_multi over all qubits of
    |0..0> is the tensor power of gate|0>; c<gate>(control=n-1,target=0) with the control clear
    is the identity, with the control set applies gate to the target
    (e.g. operator_tests.rs:51-81, 83-117, 1304-1351, 1353-1424)."""
    m, args, _ = SINGLE[name]
    angle_args = args
    out = getattr(qi.State.new_zero(n), f"{name}_multi")(list(range(n)), *angle_args)
    assert_amps(out, kron_all([m @ ZERO] * n), msg=f"{name}_multi {n}q")
    c, t = n - 1, 0
    z = qi.State.new_zero(n)
    assert_amps(getattr(z, f"c{name}_multi")([t], [c], *angle_args), vec(z), msg="control clear")
    t1 = _state(qi, basis(n, 1 << t))                    # control 0, target 1
    assert_amps(getattr(t1, f"c{name}_multi")([t], [c], *angle_args), vec(t1), msg="control clear, target 1")
    for tv, tbit in ((ZERO, 0), (ONE, 1)):
        st = _state(qi, basis(n, (1 << c) | (tbit << t)))
        exp = np.zeros(1 << n, dtype=complex)
        mv = m @ tv
        exp[1 << c] = mv[0]
        exp[(1 << c) | (1 << t)] = mv[1]
        assert_amps(getattr(st, f"c{name}_multi")([t], [c], *angle_args), exp, msg=f"c{name} {n}q control set")


def test_h_specific(qi):
    """operator_tests.rs:22-49: h|0>=|+>, h|1>=|->, h|+>=|0>, h|->=|1>; h_multi|00>=|++>;
    ch(c=0,t=1)|+1> = |01>; ch(c=1,t=0) on (|00>) x |+> unchanged."""
    S = qi.State
    assert S.new_zero(1).h(0) == S.new_plus(1)
    assert S.new_basis_n(1, 1).h(0) == S.new_minus(1)
    assert S.new_plus(1).h(0) == S.new_zero(1)
    assert S.new_minus(1).h(0) == S.new_basis_n(1, 1)
    assert S.new_zero(2).h_multi([0, 1]) == S.new_plus(2)
    st = S.new_plus(1).tensor_product(S.new_basis_n(1, 1))
    assert st.ch_multi([1], [0]) == S.new_basis_n(2, 1)
    st = S.new_basis_n(2, 0).tensor_product(S.new_plus(1))
    assert st.ch_multi([0], [1]) == st
    assert qi.Hadamard().base_qubits() == 1


def test_pauli_specific(qi):
    """operator_tests.rs:129-165, 245-280, 345-375: eigenstate facts, self-inverse, cx with 2 controls."""
    S = qi.State
    assert S.new_minus(1).x(0) == S.new_minus(1) * -1.0
    assert S.new_zero(1).x(0).x(0) == S.new_zero(1)
    assert S.new_basis_n(2, 3).cx_multi([1], [0]) == S.new_basis_n(2, 1)
    assert S.new_basis_n(3, 3).cx_multi([2], [0, 1]) == S.new_basis_n(3, 7)
    assert S.new_basis_n(2, 2).cx_multi([1], [0]) == S.new_basis_n(2, 2)
    assert S.new_zero(1).y(0) == 1j * S.new_basis_n(1, 1)
    assert S.new_basis_n(1, 1).y(0) == -1j * S.new_zero(1)
    assert S.new_plus(1).y(0) == -1j * S.new_minus(1)
    assert S.new_minus(1).y(0) == 1j * S.new_plus(1)
    assert S.new_zero(2).y_multi([0, 1]) == S.new_basis_n(2, 3) * complex(-1.0, 0.0)
    assert S.new_zero(1).y(0).y(0) == S.new_zero(1)
    exp = (S.new_zero(1) * -1j).tensor_product(S.new_basis_n(1, 1))
    assert S.new_basis_n(2, 3).cy_multi([1], [0]) == exp
    assert S.new_plus(1).z(0) == S.new_minus(1)
    assert S.new_plus(1).z(0).z(0) == S.new_plus(1)
    exp = (S.new_basis_n(1, 1) * -1.0).tensor_product(S.new_basis_n(1, 1))
    assert S.new_basis_n(2, 3).cz_multi([1], [0]) == exp
    for p in (qi.Pauli.X, qi.Pauli.Y, qi.Pauli.Z):
        assert p.base_qubits() == 1
    # y_multi on n qubits of |0..0> = i^n |1..1>  (operator_tests.rs:283-293, 316-326)
    for n in (N_PAR, N_BIG):
        out = S.new_zero(n).y_multi(list(range(n)))
        assert_amps(out, (1j ** n) * basis(n, (1 << n) - 1))


def test_pauli_to_pauli_string(qi):
    """operator_tests.rs:450-484."""
    S = qi.State
    for p, name in ((qi.Pauli.X, "x"), (qi.Pauli.Y, "y"), (qi.Pauli.Z, "z")):
        ps = p.to_pauli_string(0)
        assert len(ps) == 1 and ps.get_targets() == [0]
        for st in (S.new_zero(1), S.new_basis_n(1, 1), S.new_plus(1), S.new_minus(1)):
            assert getattr(st, name)(0) == ps.apply(st)


def test_phase_family_identities(qi):
    """operator_tests.rs:1045-1048 (s_dag.s = I), 1300-1303 (t_dag.t = I), 1141-1151
    (p(+-pi/2) = s/s_dag, p(+-pi/4) = t/t_dag), 1159-1170 (cp on |+1>, |+0>)."""
    S = qi.State
    plus = S.new_plus(1)
    assert plus.s(0).s_dag(0) == plus
    assert plus.t(0).t_dag(0) == plus
    assert plus.p(0, -PI / 2) == plus.s_dag(0)
    assert plus.p(0, PI / 2) == plus.s(0)
    assert plus.p(0, -PI / 4) == plus.t_dag(0)
    assert plus.p(0, PI / 4) == plus.t(0)
    st = S.new_plus(1).tensor_product(S.new_basis_n(1, 1))
    exp = (S2 * (S.new_zero(1) + 1j * S.new_basis_n(1, 1))).tensor_product(S.new_basis_n(1, 1))
    assert st.cp_multi([1], [0], PI / 2) == exp
    st = S.new_basis_n(2, 0).tensor_product(S.new_plus(1))
    assert st.cp_multi([0], [1], -PI / 2) == st
    assert qi.PhaseShift.new(THETA).base_qubits() == 1
    for op in (qi.PhaseS(), qi.PhaseT(), qi.PhaseSdag(), qi.PhaseTdag(), qi.Identity()):
        assert op.base_qubits() == 1


def test_rotation_specific(qi):
    """operator_tests.rs:1266-1299 (rx eigen-facts, rx_multi|00>, crx(+-pi)); 1311-1322 tensor power
    with angle pi/1.5; 1325-1351 crx with angle pi/2.2; ry/rz analogues 1428-1750."""
    S = qi.State
    c, s = math.cos(THETA / 2), math.sin(THETA / 2)
    assert_amps(S.new_plus(1).rx(0, THETA), cmath.exp(-1j * THETA / 2) * PLUS)
    assert_amps(S.new_minus(1).rx(0, THETA), cmath.exp(1j * THETA / 2) * MINUS)
    exp = np.array([c * c, -1j * c * s, -1j * c * s, -s * s])
    assert_amps(S.new_zero(2).rx_multi([0, 1], THETA), exp)
    exp = (S.new_zero(1) * -1j).tensor_product(S.new_basis_n(1, 1))
    assert S.new_basis_n(2, 3).crx_multi([1], [0], PI) == exp
    st = S.new_basis_n(2, 0).tensor_product(S.new_plus(1))
    assert st.crx_multi([0], [1], -PI) == st
    for name, mat in (("rx", M_RX), ("ry", M_RY), ("rz", M_RZ)):
        for n in (N_PAR, N_BIG):
            a = PI / 1.5
            out = getattr(S.new_zero(n), f"{name}_multi")(list(range(n)), a)
            assert_amps(out, kron_all([mat(a) @ ZERO] * n), msg=f"{name}_multi pi/1.5 {n}q")
            a = PI / 2.2
            ctrl = n - 1
            st = _state(qi, basis(n, 1 << ctrl))
            out = getattr(st, f"c{name}_multi")([0], [ctrl], a)
            exp = np.zeros(1 << n, dtype=complex)
            mv = mat(a) @ ZERO
            exp[1 << ctrl], exp[(1 << ctrl) | 1] = mv[0], mv[1]
            assert_amps(out, exp, msg=f"c{name} pi/2.2 {n}q")
    assert qi.RotateX.new(THETA).base_qubits() == 1
    assert qi.RotateY.new(THETA).base_qubits() == 1
    assert qi.RotateZ.new(THETA).base_qubits() == 1


def test_unitary2(qi):
    """operator_tests.rs:1752-1843: U = X; non-unitary matrix rejected."""
    S = qi.State
    u = [[0j, 1 + 0j], [1 + 0j, 0j]]
    assert S.new_zero(1).unitary(0, u) == S.new_basis_n(1, 1)
    assert S.new_basis_n(1, 1).unitary(0, u) == S.new_zero(1)
    assert S.new_plus(1).unitary(0, u) == S.new_plus(1)
    assert S.new_minus(1).unitary(0, u) == S.new_minus(1) * -1.0
    assert S.new_zero(2).unitary_multi([0, 1], u) == S.new_basis_n(2, 3)
    assert S.new_basis_n(2, 3).cunitary_multi([1], [0], u) == S.new_basis_n(2, 1)
    st = S.new_basis_n(2, 0).tensor_product(S.new_plus(1))
    assert st.cunitary_multi([0], [1], u) == st
    assert qi.Unitary2.new(u).base_qubits() == 1
    with pytest.raises(qi.Error) as e:
        qi.Unitary2.new([[0j, 1 + 0j], [1 + 0j, 1 + 0j]])
    assert e.value.variant == "NonUnitaryMatrix"
    for n in (N_PAR, N_BIG):
        assert S.new_zero(n).unitary_multi(list(range(n)), u) == S.new_basis_n(n, (1 << n) - 1)
        c = n - 1
        z = S.new_zero(n)
        assert z.cunitary_multi([0], [c], u) == z
        st = _state(qi, basis(n, 1 << c))
        assert_amps(st.cunitary_multi([0], [c], u), basis(n, (1 << c) | 1))


def test_ry_phase(qi):
    """operator_tests.rs:1845-1953: ry_phase(theta,phi) = p(phi) then ry(theta); special cases
    RY (phi=0), P (theta=0), H (pi/2, pi), X (pi, pi); 1955-1981: ry_phase_dag inverts ry_phase."""
    S = qi.State
    theta, phi = PI / 2.5, PI / 4.0
    for st in (S.new_zero(1), S.new_basis_n(1, 1), S.new_plus(1), S.new_minus(1)):
        assert st.ry_phase(0, theta, phi) == st.p(0, phi).ry(0, theta)
    threes = (S.new_zero(3), S.new_basis_n(3, 1), S.new_plus(3), S.new_minus(3))
    for st in threes:
        assert st.ry_phase(0, theta, 0.0) == st.ry(0, theta)
        assert st.ry_phase(0, 0.0, phi) == st.p(0, phi)
        assert st.ry_phase(0, PI / 2, PI) == st.h(0)
        assert st.ry_phase(0, PI, PI) == st.x(0)
    theta, phi = PI / 1.5, PI / 3.25
    for st in threes:
        assert st.ry_phase(0, theta, phi).ry_phase_dag(0, theta, phi) == st
    # multi / controlled variants exist with the reference's argument order (state.rs:2072-2220)
    st = S.new_plus(3)
    assert st.ry_phase_multi([0, 2], theta, phi) == st.ry_phase(0, theta, phi).ry_phase(2, theta, phi)
    assert st.cry_phase_gates([0], [1], theta, phi).cry_phase_dag_gates([0], [1], theta, phi) == st
    assert st.ry_phase_dag_multi([1], theta, phi) == st.ry_phase_dag(1, theta, phi)


def test_cnot(qi):
    """operator_tests.rs:1983-2065: cnot(control, target) truth table, Bell pair, 11q, Toffoli as CCX."""
    S = qi.State
    assert S.new_zero(2).cnot(0, 1) == S.new_zero(2)
    assert S.new_basis_n(2, 2).cnot(0, 1) == S.new_basis_n(2, 2)
    assert S.new_basis_n(2, 1).cnot(0, 1) == S.new_basis_n(2, 3)
    assert S.new_basis_n(2, 3).cnot(0, 1) == S.new_basis_n(2, 1)
    st = S2 * (S.new_zero(2) + S.new_basis_n(2, 1))
    assert st.cnot(0, 1) == complex(S2, 0.0) * (S.new_basis_n(2, 0) + S.new_basis_n(2, 3))
    assert qi.CNOT().base_qubits() == 2
    for n in (N_PAR, N_BIG):
        c, t = n - 1, 0
        z = S.new_zero(n)
        assert z.cnot(c, t) == z
        assert_amps(_state(qi, basis(n, 1 << c)).cnot(c, t), basis(n, (1 << c) | 1))
        c2 = n - 2
        assert z.toffoli(c, c2, t) == z
        assert_amps(_state(qi, basis(n, (1 << c) | (1 << c2))).toffoli(c, c2, t),
                    basis(n, (1 << c) | (1 << c2) | 1))


def test_swap(qi):
    """operator_tests.rs:2067-2158."""
    S = qi.State
    assert S.new_zero(2).swap(0, 1) == S.new_zero(2)
    assert S.new_basis_n(2, 2).swap(0, 1) == S.new_basis_n(2, 1)
    assert S.new_basis_n(2, 1).swap(0, 1) == S.new_basis_n(2, 2)
    assert S.new_basis_n(2, 3).swap(0, 1) == S.new_basis_n(2, 3)
    st = 0.5 * (S.new_zero(2) - S.new_basis_n(2, 1) + S.new_basis_n(2, 2) - S.new_basis_n(2, 3))
    ex = 0.5 * (S.new_zero(2) + S.new_basis_n(2, 1) - S.new_basis_n(2, 2) - S.new_basis_n(2, 3))
    assert st.swap(0, 1) == ex
    assert S.new_basis_n(3, 3).cswap(1, 2, [0]) == S.new_basis_n(3, 5)
    assert S.new_basis_n(3, 4).cswap(1, 2, [0]) == S.new_basis_n(3, 4)
    assert qi.SWAP().base_qubits() == 2
    for n in (N_PAR, N_BIG):
        assert_amps(_state(qi, basis(n, 1)).swap(0, 1), basis(n, 2))
        c = n - 1
        st = _state(qi, basis(n, 1))
        assert st.cswap(0, 1, [c]) == st
        assert_amps(_state(qi, basis(n, (1 << c) | 1)).cswap(0, 1, [c]), basis(n, (1 << c) | 2))


def test_matchgate(qi):
    """operator_tests.rs:2210-2328 (theta = pi cases) and the 15-qubit block 2330-2388."""
    S = qi.State
    st = S.new_basis_n(3, 5)
    assert st.matchgate(1, PI, PI, 0.0) == S.new_basis_n(3, 3)
    assert st.matchgate(1, PI, PI, 0.0).matchgate(1, PI, PI, 0.0) == st
    th, p1, p2 = PI, PI / 2, PI / 3
    assert S.new_basis_n(2, 0).matchgate(0, th, p1, p2) == S.new_basis_n(2, 0)
    assert S.new_basis_n(2, 1).matchgate(0, th, p1, p2) == S.new_basis_n(2, 2)
    assert S.new_basis_n(2, 2).matchgate(0, th, p1, p2) == S.new_basis_n(2, 1) * complex(0.0, -1.0)
    assert S.new_basis_n(2, 3).matchgate(0, th, p1, p2) == S.new_basis_n(2, 3) * cmath.exp(1j * PI / 3)
    assert qi.Matchgate(1.0, 2.0, 3.0).base_qubits() == 2
    assert S.new_basis_n(3, 3).cmatchgate(1, PI, PI / 2, PI / 3, [0]) == S.new_basis_n(3, 5)
    assert S.new_basis_n(3, 4).cmatchgate(1, PI, PI / 2, PI / 3, [0]) == S.new_basis_n(3, 4)
    for n in (N_PAR, N_BIG):
        top, top2 = 1 << (n - 1), 1 << (n - 2)
        assert S.new_basis_n(n, 1).matchgate(0, th, p1, p2) == S.new_basis_n(n, 2)
        assert S.new_basis_n(n, 2).matchgate(0, th, p1, p2) == S.new_basis_n(n, 1) * complex(0.0, -1.0)
        c1, c2 = [n - 1], [n - 1, n - 2]
        for k in (1, 2):
            st = S.new_basis_n(n, k)
            assert st.cmatchgate(0, th, p1, p2, c1) == st
        assert S.new_basis_n(n, top + 1).cmatchgate(0, th, p1, p2, c1) == S.new_basis_n(n, top + 2)
        assert S.new_basis_n(n, top + 2).cmatchgate(0, th, p1, p2, c1) == \
            S.new_basis_n(n, top + 1) * complex(0.0, -1.0)
        st = S.new_basis_n(n, top + 1)
        assert st.cmatchgate(0, th, p1, p2, c2) == st
        assert S.new_basis_n(n, top + top2 + 2).cmatchgate(0, th, p1, p2, c2) == \
            S.new_basis_n(n, top + top2 + 1) * complex(0.0, -1.0)


def test_matchgate_general_angle_follows_cpu_path(qi):
    """The reference only tests theta = pi, which hides a CPU/OpenCL divergence: the CPU
    path (operator.rs:960-970), which is the oracle, uses e^{i phi1} for BOTH |10>-column entries."""
    th, p1, p2 = 0.7, 0.4, 1.1
    c, s = math.cos(th / 2), math.sin(th / 2)
    e1, e2 = cmath.exp(1j * p1), cmath.exp(1j * p2)
    v = np.array([0.1 + 0.2j, 0.3 - 0.1j, -0.2 + 0.5j, 0.4 + 0.1j])
    v = v / np.linalg.norm(v)
    out = _state(qi, v).matchgate(0, th, p1, p2)
    exp = np.array([v[0], c * v[1] - e1 * s * v[2], s * v[1] + e1 * c * v[2], e2 * v[3]])
    assert_amps(out, exp)


def test_toffoli(qi):
    """operator_tests.rs:2391-2490 incl. CCCX through Pauli::X.apply with three controls."""
    S = qi.State
    table = {0: 0, 1: 1, 2: 2, 3: 7, 4: 4, 5: 5, 6: 6, 7: 3}
    for k, v in table.items():
        assert S.new_basis_n(3, k).toffoli(0, 1, 2) == S.new_basis_n(3, v)
    assert qi.Toffoli().base_qubits() == 3
    for n in (N_PAR, N_BIG):
        ctrls = [n - 1, n - 2, n - 3]
        st = _state(qi, basis(n, (1 << ctrls[0]) | (1 << ctrls[1])))
        assert qi.Pauli.X.apply(st, [0], ctrls) == st
        allc = (1 << ctrls[0]) | (1 << ctrls[1]) | (1 << ctrls[2])
        assert_amps(qi.Pauli.X.apply(_state(qi, basis(n, allc)), [0], ctrls), basis(n, allc | 1))


def test_operate(qi):
    """operator_tests.rs:2492-2507 and the checks of state.rs:970-1002."""
    S = qi.State
    assert S.new_zero(1).operate(qi.Hadamard(), [0], []) == S.new_plus(1)
    assert S.new_basis_n(2, 1).operate(qi.CNOT(), [1], [0]) == S.new_basis_n(2, 3)
    with pytest.raises(qi.Error) as e:
        S.new_zero(2).operate(qi.CNOT(), [1], [])
    assert (e.value.variant, e.value.payload) == ("InvalidNumberOfQubits", (2,))


def _raises(qi, variant, payload, fn):
    with pytest.raises(qi.Error) as e:
        fn()
    assert e.value.variant == variant, e.value
    if payload is not None:
        assert tuple(e.value.payload) == tuple(payload), e.value


def test_single_qubit_gate_errors(qi):
    """operator_tests.rs:2513-2566: index 2 on a 2-qubit state -> InvalidQubitIndex(2, 2)."""
    st = qi.State.new_zero(2)
    for name in ("h", "x", "y", "z", "s", "t", "s_dag", "t_dag", "i"):
        _raises(qi, "InvalidQubitIndex", (2, 2), lambda: getattr(st, name)(2))
        _raises(qi, "InvalidQubitIndex", (2, 2), lambda: getattr(st, f"{name}_multi")([0, 2]))
    for name in ("p", "rx", "ry", "rz"):
        _raises(qi, "InvalidQubitIndex", (2, 2), lambda: getattr(st, name)(2, PI / 4))
        _raises(qi, "InvalidQubitIndex", (2, 2), lambda: getattr(st, f"{name}_multi")([0, 2], PI / 4))


def test_multi_qubit_gate_errors(qi):
    """operator_tests.rs:2568-2703."""
    st = qi.State.new_zero(3)
    bad = 3
    _raises(qi, "InvalidQubitIndex", (3, 3), lambda: st.cnot(bad, 0))
    _raises(qi, "InvalidQubitIndex", (3, 3), lambda: st.cnot(0, bad))
    _raises(qi, "InvalidQubitIndex", (3, 3), lambda: st.swap(bad, 1))
    _raises(qi, "InvalidQubitIndex", (3, 3), lambda: st.swap(0, bad))
    _raises(qi, "InvalidQubitIndex", (3, 3), lambda: st.toffoli(bad, 1, 2))
    _raises(qi, "InvalidQubitIndex", (3, 3), lambda: st.toffoli(0, bad, 2))
    _raises(qi, "InvalidQubitIndex", (3, 3), lambda: st.toffoli(0, 1, bad))
    _raises(qi, "InvalidQubitIndex", (3, 3), lambda: st.matchgate(3, PI, 0.0, 0.0))
    _raises(qi, "InvalidQubitIndex", (2, 3), lambda: st.matchgate(2, PI, 0.0, 0.0))
    _raises(qi, "OverlappingControlAndTargetQubits", (0, 0), lambda: st.cmatchgate(0, PI, 0.0, 0.0, [0]))


def test_validation_order_and_variants(qi):
    """operator.rs:214-273 (count -> target range -> control range/overlap -> duplicate targets),
    CNOT/Toffoli control-count checks (operator.rs:677-679, 1065-1072)."""
    st = qi.State.new_zero(3)
    _raises(qi, "InvalidNumberOfQubits", (2,), lambda: qi.Hadamard().apply(st, [0, 1], []))
    _raises(qi, "InvalidNumberOfQubits", (0,), lambda: qi.Hadamard().apply(st, [], []))
    _raises(qi, "InvalidQubitIndex", (7, 3), lambda: qi.Hadamard().apply(st, [7], [9]))
    _raises(qi, "InvalidQubitIndex", (9, 3), lambda: qi.Hadamard().apply(st, [1], [9]))
    _raises(qi, "OverlappingControlAndTargetQubits", (1, 1), lambda: qi.Hadamard().apply(st, [1], [0, 1]))
    _raises(qi, "InvalidQubitIndex", (1, 3), lambda: qi.SWAP().apply(st, [1, 1], []))
    _raises(qi, "InvalidNumberOfQubits", (1,), lambda: qi.SWAP().apply(st, [1], []))
    _raises(qi, "InvalidNumberOfQubits", (0,), lambda: qi.CNOT().apply(st, [1], []))
    _raises(qi, "InvalidNumberOfQubits", (2,), lambda: qi.CNOT().apply(st, [2], [0, 1]))
    _raises(qi, "InvalidNumberOfQubits", (1,), lambda: qi.Toffoli().apply(st, [2], [0]))
    _raises(qi, "InvalidNumberOfQubits", (2,), lambda: qi.Toffoli().apply(st, [2], [0, 0]))
