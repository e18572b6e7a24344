"""Port of the reference's model-Hamiltonian tests (src/tests/ising_tests.rs, heisenberg_tests.rs:107-226):
the builders are compared with hand-listed terms by applying both SumOps to a basis state, as the reference does.
Runs against the CPU oracle everywhere and against the GPU engine on a B200 (host-only builders)."""
import pytest


def _raises(qi, variant, payload, fn):
    with pytest.raises(qi.Error) as e:
        fn()
    assert e.value.variant == variant and tuple(e.value.payload) == tuple(payload), e.value


def _z(qi, coeff, *qubits):
    ps = qi.PauliString.new(complex(coeff, 0.0))
    for q in qubits:
        ps = ps.with_op(q, qi.Pauli.Z)
    return ps


def test_ising_1d(qi):
    """ising_tests.rs:13-56."""
    h, j, m = [1.0, 2.0, 3.0], [0.5, 1.0, 1.5], 0.1
    res = qi.ising_1d(h, j, m)
    assert len(res.terms) == 6
    exp = [_z(qi, -m * h[i], i) for i in range(3)] + [_z(qi, -j[i], i, (i + 1) % 3) for i in range(3)]
    st = qi.State.new_basis_n(3, 6)
    assert res.apply(st) == qi.SumOp.new(exp).apply(st)
    _raises(qi, "InvalidNumberOfInputs", (1, 2), lambda: qi.ising_1d([1.0], [0.5], m))
    # per site: coupling first, then field (ising.rs:49-66); zero coefficients are skipped
    t0, t1 = res.terms[0], res.terms[1]
    assert sorted(t0.ops()) == [0, 1] and t0.coefficient() == complex(-0.5, 0.0)
    assert sorted(t1.ops()) == [0] and t1.coefficient() == complex(-0.1, 0.0)
    assert len(qi.ising_1d([0.0, 2.0], [1.0, 0.0], m).terms) == 2
    assert len(qi.ising_1d([0.0, 0.0], [0.0, 0.0], m).terms) == 0


def test_ising_1d_uniform(qi):
    """ising_tests.rs:58-101."""
    h, j, m = 1.0, 2.0, 0.1
    res = qi.ising_1d_uniform(3, h, j, m)
    assert len(res.terms) == 6
    exp = [_z(qi, -m * h, i) for i in range(3)] + [_z(qi, -j, i, (i + 1) % 3) for i in range(3)]
    st = qi.State.new_basis_n(3, 6)
    assert res.apply(st) == qi.SumOp.new(exp).apply(st)
    _raises(qi, "InvalidNumberOfInputs", (1, 2), lambda: qi.ising_1d_uniform(1, h, j, m))
    assert len(qi.ising_1d_uniform(4, 0.0, 0.0, m).terms) == 0
    assert len(qi.ising_1d_uniform(4, 0.0, 1.0, m).terms) == 4


def test_ising_2d(qi):
    """ising_tests.rs:103-207: 3x3 lattice, site (r,c) -> qubit 3r+c, j[r][c] = (vertical, horizontal)."""
    h = [[1.0, 2.0, 3.0], [4.0, 5.0, 6.0], [7.0, 8.0, 9.0]]
    j = [[[0.5, 1.0], [1.5, 2.0], [2.5, 3.0]], [[3.5, 4.0], [4.5, 5.0], [5.5, 6.0]], [[6.5, 7.0], [7.5, 8.0], [8.5, 9.0]]]
    m = 0.1
    res = qi.ising_2d(h, j, m)
    assert len(res.terms) == 27
    exp = []
    for r in range(3):
        for c in range(3):
            q = 3 * r + c
            exp.append(_z(qi, -0.1 * h[r][c], q))
            exp.append(_z(qi, -j[r][c][0], q, 3 * ((r + 1) % 3) + c))
            exp.append(_z(qi, -j[r][c][1], q, 3 * r + (c + 1) % 3))
    st = qi.State.new_basis_n(9, 511)
    assert res.apply(st) == qi.SumOp.new(exp).apply(st)
    # the hand-listed spot checks of the reference: Z0Z3 = -0.5, Z0Z1 = -1.0, Z8Z2 (PBC) = -8.5, Z8Z6 (PBC) = -9.0
    got = {(tuple(sorted(t.ops())), t.coefficient()) for t in res.terms}
    for qs, cf in (((0, 3), -0.5), ((0, 1), -1.0), ((2, 8), -8.5), ((6, 8), -9.0), ((1, 4), -1.5)):
        assert (qs, complex(cf, 0.0)) in got
    _raises(qi, "InvalidNumberOfInputs", (1, 2), lambda: qi.ising_2d([[1.0]], [[[0.5, 1.0]]], m))


def test_ising_2d_uniform(qi):
    """ising_tests.rs:209-316."""
    h, j, m = 1.0, 2.0, 0.1
    res = qi.ising_2d_uniform(3, 3, h, j, m)
    assert len(res.terms) == 27
    exp = [_z(qi, -m * h, q) for q in range(9)]
    for r in range(3):
        for c in range(3):
            exp.append(_z(qi, -j, 3 * r + c, 3 * ((r + 1) % 3) + c))
            exp.append(_z(qi, -j, 3 * r + c, 3 * r + (c + 1) % 3))
    st = qi.State.new_basis_n(9, 511)
    assert res.apply(st) == qi.SumOp.new(exp).apply(st)
    _raises(qi, "InvalidNumberOfInputs", (1, 2), lambda: qi.ising_2d_uniform(1, 1, h, j, m))
    _raises(qi, "InvalidNumberOfInputs", (1, 2), lambda: qi.ising_2d_uniform(2, 1, h, j, m))


def test_heisenberg_2d(qi):
    """heisenberg_tests.rs:106-227: 3x3 lattice, 63 terms; field coefficient -0.5*h*mu (heisenberg.rs:143)."""
    P = qi.Pauli
    n, m, jx, jy, jz, h, mu = 3, 3, 1.0, 2.0, 3.0, 4.0, 5.0
    res = qi.heisenberg_2d(n, m, jx, jy, jz, h, mu)
    assert len(res.terms) == 63
    exp = [(-0.5 * h * mu) * qi.PauliString.new(1.0).with_op(q, P.Z) for q in range(9)]
    for jj, p in ((jx, P.X), (jy, P.Y), (jz, P.Z)):
        for r in range(3):
            for c in range(3):
                q = 3 * r + c
                exp.append((-0.5 * jj) * qi.PauliString.new(1.0).with_op(q, p).with_op(3 * ((r + 1) % 3) + c, p))
                exp.append((-0.5 * jj) * qi.PauliString.new(1.0).with_op(q, p).with_op(3 * r + (c + 1) % 3, p))
    st = qi.State.new_basis_n(9, 255)
    assert res.apply(st) == qi.SumOp.new(exp).apply(st)
    # unlike the 1-D builder, the field sign here IS pinned by the reference's test (|255> has non-zero net Z)
    assert res.terms[0].coefficient() == complex(-10.0, 0.0) and sorted(res.terms[0].ops()) == [0]
    _raises(qi, "InvalidNumberOfInputs", (1, 2), lambda: qi.heisenberg_2d(1, 2, jx, jy, jz, h, mu))
    _raises(qi, "InvalidNumberOfInputs", (1, 2), lambda: qi.heisenberg_2d(2, 1, jx, jy, jz, h, mu))
    assert len(qi.heisenberg_2d(2, 2, 0.0, 0.0, 0.0, 0.0, mu).terms) == 0
