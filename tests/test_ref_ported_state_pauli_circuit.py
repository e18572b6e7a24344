"""Port of the reference's known-answer tests for State, measurement, PauliString/SumOp, Trotter,
Heisenberg, Gate::apply and Circuit::execute (src/tests/{state,measurement,pauli_string,
time_evolution,heisenberg,gate,circuit}_tests.rs).  Each test cites what it restates.
Runs against the CPU oracle everywhere and against the GPU engine on a B200.
"""
import cmath
import math

import numpy as np
import pytest

from conftest import assert_amps, vec

PI = math.pi
S2 = 1.0 / math.sqrt(2.0)


def _raises(qi, variant, payload, fn):
    with pytest.raises(qi.Error) as e:
        fn()
    assert e.value.variant == variant, e.value
    if payload is not None:
        assert tuple(e.value.payload) == tuple(payload), e.value


# ---------------------------------------------------------------- state_tests.rs
def test_state_new(qi):
    """state_tests.rs:5-32."""
    st = qi.State.new([1.0 + 0j, 0j])
    assert st.num_qubits == 1 and np.array_equal(vec(st), [1.0 + 0j, 0j])
    _raises(qi, "InvalidNumberOfQubits", (0,), lambda: qi.State.new([]))
    _raises(qi, "StateVectorNotNormalised", (), lambda: qi.State.new([1.0 + 0j, 1.0 + 0j]))
    _raises(qi, "InvalidNumberOfQubits", (1,), lambda: qi.State.new([1.0 + 0j, 0j, 0j]))


def test_state_constructors(qi):
    """state_tests.rs:34-141 and 303-322 (Bell states)."""
    S = qi.State
    assert S.new_hartree_fock(2, 2) == S.new_basis_n(2, 3)
    assert S.new_hartree_fock(4, 6) == S.new_basis_n(6, 60)
    assert S.new_hartree_fock(1, 4) == S.new_basis_n(4, 8)
    z = S.new_zero(1)
    assert z.num_qubits == 1 and np.array_equal(vec(z), [1, 0])
    _raises(qi, "InvalidNumberOfQubits", (0,), lambda: S.new_zero(0))
    b = S.new_basis_n(2, 1)
    assert b.num_qubits == 2 and np.array_equal(vec(b), [0, 1, 0, 0])
    _raises(qi, "InvalidQubitIndex", (4, 2), lambda: S.new_basis_n(2, 4))
    _raises(qi, "InvalidNumberOfQubits", (0,), lambda: S.new_basis_n(0, 0))
    assert np.array_equal(vec(S.new_plus(1)), [S2, S2])
    _raises(qi, "InvalidNumberOfQubits", (0,), lambda: S.new_plus(0))
    assert np.array_equal(vec(S.new_minus(1)), [S2, -S2])
    _raises(qi, "InvalidNumberOfQubits", (0,), lambda: S.new_minus(0))
    n = 6
    assert S.new_ghz(n) == math.sqrt(0.5) * (S.new_basis_n(n, (1 << n) - 1) + S.new_zero(n))
    _raises(qi, "InvalidNumberOfQubits", (0,), lambda: S.new_ghz(0))
    a = math.sqrt(0.5)
    assert np.array_equal(vec(S.new_phi_plus()), [a, 0, 0, a])
    assert np.array_equal(vec(S.new_phi_minus()), [a, 0, 0, -a])
    assert np.array_equal(vec(S.new_psi_plus()), [0, a, a, 0])
    assert np.array_equal(vec(S.new_psi_minus()), [0, a, -a, 0])
    assert S.new_phi_plus().num_qubits == 2
    # new_minus sign = parity of the index (state.rs:262-270), checked at a parallel-path size
    m = vec(S.new_minus(7))
    par = np.array([bin(i).count("1") & 1 for i in range(128)])
    assert np.array_equal(m, np.where(par == 0, 1.0, -1.0) / math.sqrt(128.0))


def test_state_normalise_probability(qi):
    """state_tests.rs:143-183 (note the literal State{..} with num_qubits that does not match)."""
    st = qi.State([3.0 + 0j, 4.0 + 0j], 2)
    assert np.array_equal(vec(st.normalise()), [0.6 + 0j, 0.8 + 0j])
    _raises(qi, "ZeroNorm", (), lambda: qi.State([0j, 0j], 2).normalise())
    assert abs(qi.State.new_plus(1).probability(0) - 0.5) < 2.3e-16
    _raises(qi, "InvalidQubitIndex", (2, 1), lambda: qi.State.new_plus(1).probability(2))
    assert qi.State.new_plus(1).amplitude(1) == complex(S2, 0.0)
    _raises(qi, "InvalidQubitIndex", (2, 1), lambda: qi.State.new_plus(1).amplitude(2))


def test_state_tensor_product_and_conj(qi):
    """state_tests.rs:185-221: a.tensor_product(b) puts `a` in the HIGH bits."""
    S = qi.State
    exp = 0.5 * (S.new_basis_n(2, 0) - S.new_basis_n(2, 1) + S.new_basis_n(2, 2) - S.new_basis_n(2, 3))
    assert S.new_plus(1).tensor_product(S.new_minus(1)) == exp
    assert S.new_zero(1).tensor_product(S.new_basis_n(1, 1)) == S.new_basis_n(2, 1)
    st = qi.State([1 + 1j, -1j], 1)
    assert st.conj() == qi.State([1 - 1j, 1j], 1)


def test_state_fs_metrics(qi):
    """state_tests.rs:223-300."""
    S = qi.State
    p2 = S.new_plus(2)
    assert abs(p2.fs_dist(p2)) < 2.3e-16
    s1, s2 = qi.State([1 + 0j, 1j], 2), qi.State([0j, 1 + 0j], 2)
    assert abs(s1.fs_dist(s2) - PI / 4) < 2.3e-16
    _raises(qi, "ZeroNorm", (), lambda: qi.State([], 0).fs_dist(qi.State([], 0)))
    _raises(qi, "InvalidNumberOfQubits", (2,), lambda: S.new_plus(2).fs_dist(S.new_plus(1)))
    _raises(qi, "ZeroNorm", (), lambda: qi.State([0j, 0j], 2).fs_dist(S.new_basis_n(2, 1)))
    assert abs(qi.State([1 + 0j, 0j], 2).fs_fidelity(qi.State([0j, 1 + 0j], 2))) < 2.3e-16
    assert abs(p2.fs_fidelity(p2) - 1.0) < 2.3e-16
    f = S.new_plus(1).fs_fidelity(S.new_basis_n(1, 1))
    assert abs(f - 0.5) < 2.3e-16 and abs(f - math.cos(PI / 4) ** 2) < 2.3e-16


# ---------------------------------------------------------------- measurement_tests.rs
SEEDS = range(24)   # enough draws to visit every outcome of the 2-qubit cases below


def test_measure_1_qubit_computational(qi):
    """measurement_tests.rs:4-55: outcome-conditional collapsed state."""
    MB = qi.MeasurementBasis
    st = qi.State.new([S2, 0, 0, S2])
    seen = set()
    for seed in SEEDS:
        r = st.measure(MB.Computational, [0], seed=seed)
        assert r.get_basis() == MB.Computational and r.get_indices() == [0]
        o = r.get_outcomes()[0]
        seen.add(o)
        assert np.array_equal(vec(r.get_new_state()), [1, 0, 0, 0] if o == 0 else [0, 0, 0, 1])
        assert r.get_new_state().num_qubits == 2
    assert seen == {0, 1}
    a = 1.0 / math.sqrt(3.0)
    st = qi.State.new([a, a, 0, a])
    for seed in SEEDS:
        r = st.measure(MB.Computational, [0], seed=seed)
        o = r.get_outcomes()[0]
        exp = [1, 0, 0, 0] if o == 0 else [0, S2, 0, S2]
        assert r.get_new_state() == qi.State.new(exp)
        assert_amps(r.get_new_state(), exp)


def test_measure_all_qubits_computational(qi):
    """measurement_tests.rs:57-80 and 119-145: empty qubit list = all qubits; outcomes[j] = bit j."""
    MB = qi.MeasurementBasis
    st = qi.State.new([0.5, 0.5, 0.5, 0.5])
    seen = set()
    for seed in SEEDS:
        r = st.measure(MB.Computational, [], seed=seed)
        o0, o1 = r.get_outcomes()
        seen.add((o1, o0))
        exp = np.zeros(4)
        exp[(o1 << 1) | o0] = 1.0
        assert np.array_equal(vec(r.get_new_state()), exp)
        assert r.get_indices() == [0, 1]
    assert len(seen) == 4
    rs = st.measure_n(MB.Computational, [], 5, seed=3)
    assert len(rs) == 5 and rs[0].get_basis() == MB.Computational
    for r in rs:
        o0, o1 = r.get_outcomes()
        exp = np.zeros(4)
        exp[(o1 << 1) | o0] = 1.0
        assert np.array_equal(vec(r.get_new_state()), exp)


def test_measure_n_1_qubit(qi):
    """measurement_tests.rs:82-117."""
    MB = qi.MeasurementBasis
    st = qi.State.new([S2, 0, 0, S2])
    rs = st.measure_n(MB.Computational, [0], 5, seed=11)
    assert len(rs) == 5 and rs[0].get_indices() == [0]
    for r in rs:
        o = r.get_outcomes()[0]
        assert np.array_equal(vec(r.get_new_state()), [1, 0, 0, 0] if o == 0 else [0, 0, 0, 1])


def test_measure_errors(qi):
    """measurement_tests.rs:99-116 and 188-211."""
    MB = qi.MeasurementBasis
    st = qi.State.new([0.5, 0.5, 0.5, 0.5])
    _raises(qi, "InvalidQubitIndex", (3, 2), lambda: st.measure(MB.Computational, [3]))
    _raises(qi, "InvalidNumberOfQubits", (2,), lambda: st.measure(MB.Computational, [0, 1, 2]))
    _raises(qi, "InvalidQubitIndex", (3, 2), lambda: st.measure_n(MB.Computational, [3], 5))
    _raises(qi, "InvalidNumberOfQubits", (2,), lambda: st.measure_n(MB.Computational, [0, 1, 2], 5))
    _raises(qi, "InvalidNumberOfMeasurements", (0,), lambda: st.measure_n(MB.Computational, [0], 0))


def test_measure_x_basis(qi):
    """measurement_tests.rs:213-268."""
    S, MB = qi.State, qi.MeasurementBasis
    st = S.new_plus(1).tensor_product(S.new_zero(1))    # |+0>: qubit1 = |+>, qubit0 = |0>
    for seed in SEEDS:
        r = st.measure(MB.X, [1], seed=seed)
        assert r.get_basis() == MB.X and r.get_indices() == [1]
        assert r.get_outcomes()[0] == 0 and r.get_new_state() == st
        r = st.measure(MB.X, [0], seed=seed)
        o = r.get_outcomes()[0]
        exp = S.new_plus(2) if o == 0 else S.new_plus(1).tensor_product(S.new_minus(1))
        assert r.get_new_state() == exp
        r = st.measure(MB.X, [], seed=seed)
        assert r.get_indices() == [0, 1]
        o = tuple(r.get_outcomes())
        assert o in ((0, 0), (1, 0))
        exp = S.new_plus(2) if o == (0, 0) else S.new_plus(1).tensor_product(S.new_minus(1))
        assert r.get_new_state() == exp


def test_measure_y_basis(qi):
    """measurement_tests.rs:270-407."""
    S, MB = qi.State, qi.MeasurementBasis
    st = S.new_plus(1).tensor_product(S.new_zero(1))
    b = [S.new_basis_n(2, k) for k in range(4)]
    e0 = 0.5 * (b[0] + 1j * b[1] + b[2] + 1j * b[3])
    e1 = 0.5 * (b[0] - 1j * b[1] + b[2] - 1j * b[3])
    f0 = 0.5 * ((1 - 1j) * b[0] + 0.0 * b[1] + (1 + 1j) * b[2] + 0.0 * b[3])
    f1 = 0.5 * ((1 + 1j) * b[0] + 0.0 * b[1] + (1 - 1j) * b[2] + 0.0 * b[3])
    d = 2.0 * math.sqrt(2.0)
    all_exp = {
        (0, 0): [(1 - 1j) / d, (1 + 1j) / d, (1 + 1j) / d, (-1 + 1j) / d],
        (1, 0): [(1 - 1j) / d, (-1 - 1j) / d, (1 + 1j) / d, (1 - 1j) / d],
        (0, 1): [(1 + 1j) / d, (-1 + 1j) / d, (1 - 1j) / d, (1 + 1j) / d],
        (1, 1): [(1 + 1j) / d, (1 - 1j) / d, (1 - 1j) / d, (-1 - 1j) / d],
    }
    seen = set()
    for seed in SEEDS:
        r = st.measure(MB.Y, [0], seed=seed)
        assert r.get_basis() == MB.Y and r.get_indices() == [0]
        assert r.get_new_state() == (e0 if r.get_outcomes()[0] == 0 else e1)
        r = st.measure(MB.Y, [1], seed=seed)
        assert r.get_new_state() == (f0 if r.get_outcomes()[0] == 0 else f1)
        r = st.measure(MB.Y, [], seed=seed)
        o = tuple(r.get_outcomes())
        seen.add(o)
        assert r.get_new_state() == S.new(all_exp[o])
    assert len(seen) == 4


def test_measure_custom_basis(qi):
    """state.rs:706-728: Custom(U) applies U before and U^dagger after.  U = H reproduces the
    X-basis facts of measurement_tests.rs:213-246."""
    S, MB = qi.State, qi.MeasurementBasis
    h = [[S2 + 0j, S2 + 0j], [S2 + 0j, -S2 + 0j]]
    st = S.new_plus(1).tensor_product(S.new_zero(1))
    for seed in SEEDS:
        r = st.measure(MB.Custom(h), [0], seed=seed)
        o = r.get_outcomes()[0]
        exp = S.new_plus(2) if o == 0 else S.new_plus(1).tensor_product(S.new_minus(1))
        assert r.get_new_state() == exp
        assert r.get_basis() == MB.Custom(h)


# ---------------------------------------------------------------- pauli_string_tests.rs
def _xy(qi, coeff=complex(2.0, 2.0)):
    ps = qi.PauliString.new(coeff)
    ps.add_op(0, qi.Pauli.X)
    ps.add_op(1, qi.Pauli.Y)
    return ps


def test_pauli_string_container(qi):
    """pauli_string_tests.rs:12-76, 231-243, 306-316."""
    P = qi.Pauli
    ps = qi.PauliString.new(complex(1.0, 0.0))
    assert ps.coefficient() == complex(1.0, 0.0) and len(ps.ops()) == 0 and ps.len() == 0
    ps.add_op(0, P.X)
    assert ps.len() == 1
    ps.add_op(1, P.Y)
    assert ps.len() == 2 and ps.ops()[0] is P.X and ps.ops()[1] is P.Y
    with pytest.raises(Exception):
        ps.add_op(0, P.Y)          # Rust: panic on duplicate qubit (pauli_string.rs:66-70)
    with pytest.raises(Exception):
        qi.PauliString.new(1.0).with_op(0, P.X).with_op(0, P.Y)
    w = qi.PauliString.with_ops(complex(1.0, 0.0), {0: P.X, 1: P.Y})
    assert w.ops() == {0: P.X, 1: P.Y}
    hc = _xy(qi).hermitian_conjugate()
    assert hc.coefficient() == complex(2.0, -2.0) and hc.ops() == {0: P.X, 1: P.Y}
    t = qi.PauliString.new(1.0).with_op(2, P.X).with_op(0, P.Y).with_op(1, P.Z)
    assert sorted(t.get_targets()) == [0, 1, 2]
    gates = t.to_gates()
    assert len(gates) == 3
    assert sorted((g.get_target_qubits()[0], repr(g.op)) for g in gates) == \
        [(0, "Pauli.Y"), (1, "Pauli.Z"), (2, "Pauli.X")]


def test_pauli_string_apply(qi):
    """pauli_string_tests.rs:78-140."""
    coeff = complex(2.0, 2.0)
    st = qi.State.new_basis_n(2, 3)
    assert _xy(qi).apply(st) == st.x(0).y(1) * coeff
    assert_amps(_xy(qi).apply(st), vec(st.x(0).y(1)) * coeff)
    assert qi.PauliString.new(coeff).apply(st) == st * coeff
    with pytest.raises(qi.Error):
        _xy(qi).apply(qi.State.new_basis_n(1, 1))
    assert _xy(qi).apply_normalised(st) == st.x(0).y(1)
    assert qi.PauliString.new(coeff).apply_normalised(st) == st


def test_pauli_string_apply_exp(qi):
    """pauli_string_tests.rs:142-237: exp(alpha P) psi = cosh(alpha) psi + sinh(alpha) P psi."""
    coeff = complex(2.0, 2.0)
    st = qi.State.new_basis_n(2, 3)
    exp = st * cmath.cosh(coeff) + st.x(0).y(1) * cmath.sinh(coeff)
    assert _xy(qi).apply_exp(st) == exp
    assert_amps(_xy(qi).apply_exp(st), vec(st) * cmath.cosh(coeff) + vec(st.x(0).y(1)) * cmath.sinh(coeff),
                tol=1e-11)   # |cosh(2+2i)| ~ 3.8: absolute bar scaled by the amplitude size
    assert qi.PauliString.new(coeff).apply_exp(st) == st * cmath.exp(coeff)
    with pytest.raises(qi.Error):
        _xy(qi).apply_exp(qi.State.new_basis_n(1, 1))
    a = coeff * 0.5
    assert _xy(qi).apply_exp_factor(st, complex(0.5, 0.0)) == st * cmath.cosh(a) + st.x(0).y(1) * cmath.sinh(a)
    assert qi.PauliString.new(coeff).apply_exp_factor(st, complex(0.5, 0.0)) == st * cmath.exp(a)
    with pytest.raises(qi.Error):
        _xy(qi).apply_exp_factor(qi.State.new_basis_n(1, 1), complex(0.5, 0.0))


def test_pauli_string_exp_neg_i_dt_golden(qi):
    """pauli_string_tests.rs:408-461 golden vector [1/sqrt2, 0, -i/sqrt2, 0]."""
    P = qi.Pauli
    ps = qi.PauliString.new(complex(0.5 * PI, 0.0)).with_op(0, P.Z).with_op(1, P.X)
    st = qi.State.new_zero(2)
    out = ps.apply_exp_neg_i_dt(st, 0.5)
    assert out == qi.State.new([S2, 0, -1j * S2, 0])
    assert_amps(out, [S2, 0, -1j * S2, 0])
    assert ps.apply_exp_neg_i_dt(st, 0.0) == st
    with pytest.raises(qi.Error):
        ps.apply_exp_neg_i_dt(qi.State.new_zero(1), 0.5)
    bad = qi.PauliString.new(complex(1.0, 1.0)).with_op(0, P.Z).with_op(1, P.X)
    _raises(qi, "InvalidPauliStringCoefficient", None, lambda: bad.apply_exp_neg_i_dt(qi.State.new_zero(2), 0.5))


def test_sumop(qi):
    """pauli_string_tests.rs:239-364: container, apply, empty apply, expectation golden (-4)."""
    P = qi.Pauli
    terms = [qi.PauliString.new(1.0), qi.PauliString.new(2.0)]
    so = qi.SumOp.new(terms)
    assert len(so.terms) == 2 and so.num_terms() == 2
    so.add_term(qi.PauliString.new(3.0))
    assert so.num_terms() == 3
    p1 = qi.PauliString.new(2.0).with_op(0, P.X)
    p2 = qi.PauliString.new(3.0).with_op(1, P.Y)
    p3 = qi.PauliString.new(4.0).with_op(1, P.Z)
    st = qi.State.new_basis_n(2, 3)
    assert qi.SumOp.new([p1, p2]).apply(st) == 2.0 * st.x(0) + 3.0 * st.y(1)
    with pytest.raises(qi.Error):
        qi.SumOp.new([p1, p2]).apply(qi.State.new_basis_n(1, 1))
    assert qi.SumOp.new([]).apply(st) == st * 0.0
    assert qi.SumOp.new([p1, p2, p3]).expectation_value(st) == complex(-4.0, 0.0)
    assert qi.SumOp.new([]).expectation_value(st) == 0j


# ---------------------------------------------------------------- time_evolution_tests.rs
def _h_xy(qi):
    ps1 = qi.PauliString.new(1.0).with_op(0, qi.Pauli.X)
    ps2 = qi.PauliString.new(1.0).with_op(1, qi.Pauli.Y)
    return ps1, ps2, qi.SumOp.new([ps1, ps2])


def test_trotter_steps(qi):
    """time_evolution_tests.rs:12-107, 147-207: Trotter == the manual apply_exp_factor sequence."""
    ps1, ps2, h = _h_xy(qi)
    init = qi.State.new_basis_n(2, 2)
    dt = 0.1
    exp = ps2.apply_exp_factor(ps1.apply_exp_factor(init, complex(0.0, -dt)), complex(0.0, -dt))
    assert qi.first_order_trotter_step(h, init, dt) == exp
    assert_amps(qi.first_order_trotter_step(h, init, dt), vec(exp))
    e = ps1.apply_exp_factor(init, complex(0.0, -dt / 2))
    e = ps2.apply_exp_factor(e, complex(0.0, -dt))
    e = ps1.apply_exp_factor(e, complex(0.0, -dt / 2))
    assert qi.second_order_trotter_step(h, init, dt) == e
    e1 = init
    for _ in range(3):
        e1 = ps2.apply_exp_factor(ps1.apply_exp_factor(e1, complex(0.0, -dt)), complex(0.0, -dt))
    out = qi.trotter_evolve_state(h, init, dt, 3, qi.TrotterOrder.First)
    assert out == e1
    assert_amps(out, vec(e1))
    e2 = init
    for _ in range(3):
        e2 = ps1.apply_exp_factor(e2, complex(0.0, -dt / 2))
        e2 = ps2.apply_exp_factor(e2, complex(0.0, -dt))
        e2 = ps1.apply_exp_factor(e2, complex(0.0, -dt / 2))
    assert qi.trotter_evolve_state(h, init, dt, 3, qi.TrotterOrder.Second) == e2


def test_trotter_errors(qi):
    """time_evolution_tests.rs:42-78, 109-145, 209-243."""
    empty = qi.SumOp.new([])
    init = qi.State.new_basis_n(2, 0)
    for fn in (lambda: qi.first_order_trotter_step(empty, init, 0.1),
               lambda: qi.second_order_trotter_step(empty, init, 0.1),
               lambda: qi.trotter_evolve_state(empty, init, 0.1, 3, qi.TrotterOrder.First)):
        with pytest.raises(qi.Error):
            fn()
    bad = qi.SumOp.new([qi.PauliString.new(1.0).with_op(0, qi.Pauli.X),
                        qi.PauliString.new(1.0).with_op(2, qi.Pauli.Y)])
    init = qi.State.new_basis_n(2, 2)
    for fn in (lambda: qi.first_order_trotter_step(bad, init, 0.1),
               lambda: qi.second_order_trotter_step(bad, init, 0.1),
               lambda: qi.trotter_evolve_state(bad, init, 0.1, 3, qi.TrotterOrder.First)):
        with pytest.raises(qi.Error):
            fn()


# ---------------------------------------------------------------- heisenberg_tests.rs
def test_heisenberg_1d(qi):
    """heisenberg_tests.rs:11-103: term SET vs hand-listed terms, compared by applying to |0110>."""
    P = qi.Pauli
    n, jx, jy, jz, h, mu = 4, 1.0, 2.0, 3.0, 4.0, 5.0
    res = qi.heisenberg_1d(n, jx, jy, jz, h, mu)
    assert len(res.terms) == 16
    exp_terms = [(-0.5 * h * mu) * qi.PauliString.new(1.0).with_op(i, P.Z) for i in range(4)]
    for j, p in ((jx, P.X), (jy, P.Y), (jz, P.Z)):
        for i in range(4):
            exp_terms.append((-0.5 * j) * qi.PauliString.new(1.0).with_op(i, p).with_op((i + 1) % 4, p))
    st = qi.State.new_basis_n(4, 6)
    assert res.apply(st) == qi.SumOp.new(exp_terms).apply(st)
    _raises(qi, "InvalidNumberOfInputs", (1, 2), lambda: qi.heisenberg_1d(1, jx, jy, jz, h, mu))


def test_heisenberg_1d_term_order_and_field_sign(qi):
    """Unpinned by the reference's tests: follow the code.  Per site XX, YY, ZZ, Z(i)
    (heisenberg.rs:68-96); field coefficient -mu * (-h/2) = +mu*h/2 (heisenberg.rs:49)."""
    res = qi.heisenberg_1d(3, 1.0, 2.0, 3.0, 0.5, 0.1)
    kinds = []
    for t in res.terms:
        ops = t.ops()
        kinds.append((tuple(sorted(ops)), "".join(sorted(repr(o)[-1] for o in ops.values())), t.coefficient()))
    assert kinds[0] == ((0, 1), "XX", complex(-0.5, 0.0))
    assert kinds[1] == ((0, 1), "YY", complex(-1.0, 0.0))
    assert kinds[2] == ((0, 1), "ZZ", complex(-1.5, 0.0))
    assert kinds[3][0] == (0,) and kinds[3][1] == "Z" and kinds[3][2] == complex(0.025, 0.0)
    assert kinds[8][0] == (0, 2)      # periodic neighbour of the last site


# ---------------------------------------------------------------- gate_tests.rs / circuit_tests.rs
def test_gate_apply(qi):
    """gate_tests.rs:92-187 (Parametric gates resolve to Operator gates first, gate.rs:107-114)."""
    S, G, P = qi.State, qi.Gate, qi.Pauli
    assert G.Operator(qi.Hadamard(), [0], []).apply(S.new_zero(1)) == S.new_plus(1)
    out = G.Measurement(qi.MeasurementBasis.Computational, [0]).apply(S.new_plus(1))
    assert out == S.new_zero(1) or out == S.new_basis_n(1, 1)
    assert G.Operator(qi.RotateY(PI / 2), [0], []).apply(S.new_zero(1)) == S.new_plus(1)
    g = G.PauliString(qi.PauliString.new(2.0).with_op(0, P.X).with_op(1, P.X))
    assert g.apply(S.new_zero(2)) == S.new_basis_n(2, 3)        # coefficient dropped, then normalised
    g = G.PauliTimeEvolution(qi.PauliString.new(0.5 * PI).with_op(0, P.Z).with_op(1, P.X), 0.5)
    assert g.apply(S.new_zero(2)) == S.new([S2, 0, -1j * S2, 0])
    assert g.get_target_qubits() == [0, 1]
    assert G.Operator(qi.CNOT(), [0, 1], [2]).get_control_qubits() == [2]
    assert G.Measurement(qi.MeasurementBasis.Computational, [0, 2]).get_control_qubits() is None


def test_circuit(qi):
    """circuit_tests.rs:8-132."""
    G, S = qi.Gate, qi.State
    c = qi.Circuit(2)
    assert c.num_qubits == 2 and c.gates == []
    c = qi.Circuit.with_gates([G.h_gate(1), G.cnot_gate(0, 1)], 2)
    assert c.get_num_qubits() == 2 and len(c.get_gates()) == 2
    _raises(qi, "InvalidQubitIndex", (3, 2), lambda: qi.Circuit.with_gates([G.h_gate(1), G.cnot_gate(0, 3)], 2))
    c = qi.Circuit(2)
    c.add_gate(G.h_gate(1))
    assert len(c.gates) == 1
    _raises(qi, "InvalidQubitIndex", (3, 2), lambda: c.add_gate(G.cnot_gate(0, 3)))
    _raises(qi, "InvalidQubitIndex", (3, 2), lambda: c.add_gates([G.h_gate(1), G.cnot_gate(0, 3)]))
    c = qi.Circuit.with_gates(G.h_multi_gate([0, 1]), 2)
    assert c.execute(S.new_zero(2)) == S.new_plus(2)
    c = qi.Circuit.with_gates([G.h_gate(0), G.cnot_gate(0, 1)], 2)
    _raises(qi, "InvalidNumberOfQubits", (1,), lambda: c.execute(S.new_zero(1)))
    assert len(c.trace_execution(S.new_zero(2))) == 3
    _raises(qi, "InvalidNumberOfQubits", (1,), lambda: c.trace_execution(S.new_zero(1)))


def test_circuit_builder_argument_order(qi):
    """State::cnot(control, target) but CircuitBuilder::cnot_gate(target, control)
    (circuit.rs:1071); toffoli_gate(control1, control2, target) (circuit.rs:1118-1123)."""
    S = qi.State
    c = qi.CircuitBuilder(3).x_gate(0).cnot_gate(1, 0).build()
    assert c.execute(S.new_zero(3)) == S.new_basis_n(3, 3)
    c = qi.CircuitBuilder(3).x_gates([0, 1]).toffoli_gate(0, 1, 2).build()
    assert c.execute(S.new_zero(3)) == S.new_basis_n(3, 7)
    c = qi.CircuitBuilder(2).h_gates([0, 1]).cp_gates([1], [0], PI).build()
    assert_amps(c.execute(S.new_zero(2)), [0.5, 0.5, 0.5, -0.5])


# ---- ChainableState (state.rs:2375-2684): gate methods on Result<State, Error> ---------------------------------------------
def test_chainable_state_forwards_while_ok_and_short_circuits_on_the_first_error(qi):
    """`impl ChainableState for Result<State, Error>` is `self.and_then(|state| state.method(..))` for every gate method
    (state.rs:2611-2620): the chain evaluates to the state while every link succeeds and to the FIRST error otherwise."""
    from quant_iron_b200.chain import chain
    ok = chain(qi.State.new_zero(2)).h(0).cnot(0, 1)
    assert ok.is_ok() and not ok.is_err()
    bell = ok.unwrap()
    v = np.asarray(bell.state_vector)
    assert np.allclose(v, [1 / math.sqrt(2), 0, 0, 1 / math.sqrt(2)], atol=1e-15)
    direct = qi.State.new_zero(2).h(0).cnot(0, 1)
    assert np.array_equal(np.asarray(direct.state_vector), v)
    bad = chain(qi.State.new_zero(2)).h(0).x(5).cnot(0, 1).h(7)          # x(5): InvalidQubitIndex(5, 2); the later links never run
    assert bad.is_err() and bad.ok() is None
    assert bad.unwrap_err().variant == "InvalidQubitIndex" and tuple(bad.unwrap_err().payload)[:2] == (5, 2)
    with pytest.raises(Exception) as e:
        bad.unwrap()
    assert e.value.variant == "InvalidQubitIndex"
    # and_then with a user function, operate through the chain
    rot = chain(qi.State.new_zero(1)).operate(qi.Hadamard(), [0], []).and_then(lambda s: s.rz(0, 0.25))
    assert rot.is_ok() and abs(abs(rot.unwrap().amplitude(0)) - 1 / math.sqrt(2)) < 1e-15
