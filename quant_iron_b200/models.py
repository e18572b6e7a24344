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
