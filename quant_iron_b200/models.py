"""Model Hamiltonians (src/models/heisenberg.rs): host-side term builders, no kernel work."""
from __future__ import annotations

from .errors import Error
from .operators import Pauli
from .pauli import PauliString, SumOp


def heisenberg_1d(n: int, jx: float, jy: float, jz: float, h: float, mu: float) -> SumOp:
    """models/heisenberg.rs:28-102.  Term order per site i: XX(i,i+1), YY, ZZ, Z(i), periodic;
    coupling coefficients -J/2; the field coefficient follows the code (heisenberg.rs:49):
    -mu * (-h/2) = +mu*h/2 (the doc comment says otherwise; the code is ground truth)."""
    if n < 2:
        raise Error("InvalidNumberOfInputs", n, 2)
    if jx == 0.0 and jy == 0.0 and jz == 0.0 and h == 0.0:
        return SumOp([])
    cx, cy, cz = complex(-0.5 * jx, 0.0), complex(-0.5 * jy, 0.0), complex(-0.5 * jz, 0.0)
    field = complex(-mu * (-0.5 * h), -mu * 0.0)
    terms = []
    for i in range(n):
        j = (i + 1) % n
        if jx != 0.0:
            terms.append(PauliString(cx).with_op(i, Pauli.X).with_op(j, Pauli.X))
        if jy != 0.0:
            terms.append(PauliString(cy).with_op(i, Pauli.Y).with_op(j, Pauli.Y))
        if jz != 0.0:
            terms.append(PauliString(cz).with_op(i, Pauli.Z).with_op(j, Pauli.Z))
        if h != 0.0:
            terms.append(PauliString(field).with_op(i, Pauli.Z))
    return SumOp(terms)


def heisenberg_2d(n_rows: int, m_cols: int, jx: float, jy: float, jz: float, h_field: float, mu: float) -> SumOp:
    """models/heisenberg.rs:122-220.  Site (r, c) -> qubit r*m_cols + c, periodic in both directions.  Per site:
    field Z, then the vertical bond (XX, YY, ZZ), then the horizontal bond (XX, YY, ZZ); couplings -J/2; the
    field coefficient follows the code (heisenberg.rs:143): mu * (-h/2) — the opposite sign of heisenberg_1d."""
    if n_rows < 2:
        raise Error("InvalidNumberOfInputs", n_rows, 2)
    if m_cols < 2:
        raise Error("InvalidNumberOfInputs", m_cols, 2)
    if jx == 0.0 and jy == 0.0 and jz == 0.0 and h_field == 0.0:
        return SumOp([])
    couplings = [(complex(-0.5 * jx, 0.0), Pauli.X, jx != 0.0), (complex(-0.5 * jy, 0.0), Pauli.Y, jy != 0.0),
                 (complex(-0.5 * jz, 0.0), Pauli.Z, jz != 0.0)]
    field = complex(mu * (-0.5 * h_field), mu * 0.0)
    terms = []
    for site in range(n_rows * m_cols):
        r, c = divmod(site, m_cols)
        if h_field != 0.0:
            terms.append(PauliString(field).with_op(site, Pauli.Z))
        for neighbour in (((r + 1) % n_rows) * m_cols + c, r * m_cols + (c + 1) % m_cols):
            for coeff, p, on in couplings:
                if on:
                    terms.append(PauliString(coeff).with_op(site, p).with_op(neighbour, p))
    return SumOp(terms)


def ising_1d(h, j, mu: float) -> SumOp:
    """models/ising.rs:27-75: H = -sum_i J_i Z_i Z_{i+1} - mu sum_i h_i Z_i, periodic; per site the coupling term
    comes first, then the field term; zero coefficients are skipped.  `h`, `j`: sequences of equal length N."""
    h, j = [float(x) for x in h], [float(x) for x in j]
    n = len(h)
    if len(j) != n:
        raise Error("MismatchedNumberOfParameters", n, len(j))     # the reference enforces this through [f64; N]
    if n < 2:
        raise Error("InvalidNumberOfInputs", n, 2)
    if all(x == 0.0 for x in h) and all(x == 0.0 for x in j):
        return SumOp([])
    terms = []
    for i in range(n):
        if j[i] != 0.0:
            terms.append(PauliString(complex(j[i] * -1.0, 0.0)).with_op(i, Pauli.Z).with_op((i + 1) % n, Pauli.Z))
        if h[i] != 0.0:
            terms.append(PauliString(complex(-1.0 * mu * h[i], 0.0)).with_op(i, Pauli.Z))
    return SumOp(terms)


def ising_1d_uniform(n: int, h: float, j: float, mu: float) -> SumOp:
    """models/ising.rs:90-139."""
    if n < 2:
        raise Error("InvalidNumberOfInputs", n, 2)
    return ising_1d([h] * n, [j] * n, mu)


def ising_2d(h, j, mu: float) -> SumOp:
    """models/ising.rs:161-244.  h[r][c]: field; j[r][c] = (vertical, horizontal) coupling of site (r, c) to
    ((r+1)%N, c) and (r, (c+1)%M).  Per site: field, vertical, horizontal; zero coefficients are skipped."""
    n = len(h)
    m = len(h[0]) if n else 0
    if n < 2:
        raise Error("InvalidNumberOfInputs", n, 2)
    if m < 2:
        raise Error("InvalidNumberOfInputs", m, 2)
    if all(float(h[r][c]) == 0.0 and float(j[r][c][0]) == 0.0 and float(j[r][c][1]) == 0.0 for r in range(n) for c in range(m)):
        return SumOp([])
    terms = []
    for site in range(n * m):
        r, c = divmod(site, m)
        hv, jv, jh = float(h[r][c]), float(j[r][c][0]), float(j[r][c][1])
        if hv != 0.0:
            terms.append(PauliString(complex(-1.0 * mu * hv, 0.0)).with_op(site, Pauli.Z))
        if jv != 0.0:
            terms.append(PauliString(complex(jv * -1.0, 0.0)).with_op(site, Pauli.Z).with_op(((r + 1) % n) * m + c, Pauli.Z))
        if jh != 0.0:
            terms.append(PauliString(complex(jh * -1.0, 0.0)).with_op(site, Pauli.Z).with_op(r * m + (c + 1) % m, Pauli.Z))
    return SumOp(terms)


def ising_2d_uniform(n: int, m: int, h: float, j: float, mu: float) -> SumOp:
    """models/ising.rs:259-324."""
    if n < 2:
        raise Error("InvalidNumberOfInputs", n, 2)
    if m < 2:
        raise Error("InvalidNumberOfInputs", m, 2)
    return ising_2d([[h] * m for _ in range(n)], [[(j, j)] * m for _ in range(n)], mu)
