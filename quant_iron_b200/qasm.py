"""OpenQASM 3.0 export of a circuit (src/compiler/{compilable,ir,qasm}.rs, `Circuit::to_qasm`, circuit.rs:204-278).

Host-only string emission: nothing here touches amplitudes or the device.  The output follows the reference's
emitter instruction for instruction, including its conventions (control prefix written before the gate name,
`U(theta, phi, lambda)` decomposition of a custom 2x2 with three decimals, one bit register per measurement gate,
`xmeasure` / `ymeasure` helper definitions in the header).
"""
from __future__ import annotations

import cmath
import math
import os
from decimal import Decimal
from typing import List, Optional, Tuple

from .errors import CompilerError
from .measurement import MeasurementBasis


def _f64(x: float) -> str:
    """Rust's `{}` for f64: shortest round-trip digits, never an exponent, no trailing `.0`."""
    if math.isnan(x):
        return "NaN"
    if math.isinf(x):
        return "inf" if x > 0 else "-inf"
    r = repr(float(x))
    if "e" in r or "E" in r:
        r = format(Decimal(r), "f")
    if r.endswith(".0"):
        r = r[:-2]
    return r


def _fmt_complex(c: complex) -> str:
    """ir.rs `format_complex`."""
    eps = 1e-9
    r, i = c.real, c.imag
    if abs(r) < eps and abs(i) < eps:
        return "0"
    if abs(i) < eps:
        return f"{r:.3f}"
    if abs(r) < eps:
        return f"{i:.3f}i"
    return f"{r:.3f} {'-' if math.copysign(1.0, i) < 0 else '+'} {abs(i):.3f}i"


def _controls(controls: List[int]) -> Tuple[str, str]:
    if not controls:
        return "", ""
    return (f"ctrl({len(controls)}) @ " + ", ".join(f"q[{c}]" for c in controls),
            "with control qubits: " + ", ".join(str(c) for c in controls))


# instruction kinds of the emitter (qasm.rs `QasmInstruction`)
_GATE, _BITREG, _MEAS, _GROUP = "gate", "bitreg", "meas", "group"


def _unitary_decl(m, target: int, controls: List[int]):
    """ir.rs `InstructionIR::Unitary`: ZYZ-style angles of U(theta, phi, lambda) up to a global phase."""
    cq, cc = _controls(controls)
    a, b, c, d = complex(m[0][0]), complex(m[0][1]), complex(m[1][0]), complex(m[1][1])
    eps = 1e-9
    if abs(1.0 - abs(a)) < eps:
        theta, phi, lam = 0.0, 0.0, cmath.phase(d) - cmath.phase(a)
    elif abs(1.0 - abs(c)) < eps:
        theta, alpha = math.pi, math.pi / 2
        phi, lam = cmath.phase(c) - alpha, cmath.phase(b) - alpha + math.pi
    else:
        if abs(a) < eps and abs(c) < eps:
            raise CompilerError("UnsupportedOperator", "Custom Unitary with zero first column")
        theta, alpha = 2.0 * math.atan2(abs(c), abs(a)), cmath.phase(a)
        phi, lam = cmath.phase(c) - alpha, cmath.phase(b) - alpha + math.pi
    comment = (f"Custom Unitary U({_fmt_complex(a)}, {_fmt_complex(b)}, {_fmt_complex(c)}, {_fmt_complex(d)}) "
               f"on qubit {target}")
    if controls:
        comment = f"{comment} {cc}"
    return (_GATE, f"{cq} U({theta:.3f},{phi:.3f},{lam:.3f}) q[{target}] // {comment}")


_SIMPLE = {  # operator class name -> (qasm mnemonic, comment)
    "Hadamard": ("h", "Hadamard gate on qubit"), "Identity": ("id", "Identity gate on qubit"),
    "PhaseS": ("s", "Phase S gate on qubit"), "PhaseT": ("t", "Phase T gate on qubit"),
    "PhaseSdag": ("sdg", "Phase S-dagger gate on qubit"), "PhaseTdag": ("tdg", "Phase T-dagger gate on qubit"),
}
_PAULI = {"Pauli.X": ("x", "Pauli-X gate on qubit"), "Pauli.Y": ("y", "Pauli-Y gate on qubit"),
          "Pauli.Z": ("z", "Pauli-Z gate on qubit")}
_ANGLE = {"PhaseShift": ("p", "Phase gate with angle"), "RotateX": ("rx", "Rotate-X gate with angle"),
          "RotateY": ("ry", "Rotate-Y gate with angle"), "RotateZ": ("rz", "Rotate-Z gate with angle")}


def _operator_instructions(op, targets: List[int], controls: List[int]):
    """compilable.rs `impl Compilable for …` followed by ir.rs `to_qasm`, for one Gate::Operator."""
    name = type(op).__name__
    out = []

    def plain(mnemonic, text, ctl):
        cq, cc = _controls(ctl)
        for t in targets:
            out.append((_GATE, f"{cq} {mnemonic} q[{t}] // {text} {t} {cc}"))

    if name in _SIMPLE:
        plain(*_SIMPLE[name], controls)
    elif repr(op) in _PAULI:
        plain(*_PAULI[repr(op)], controls)
    elif name == "CNOT":                       # only the first control is used (compilable.rs:133-146)
        plain(*_PAULI["Pauli.X"], controls[:1])
    elif name == "Toffoli":
        plain(*_PAULI["Pauli.X"], controls)
    elif name in _ANGLE:
        mnemonic, text = _ANGLE[name]
        cq, cc = _controls(controls)
        ang = _f64(op.angle)
        for t in targets:
            out.append((_GATE, f"{cq} {mnemonic}({ang}) q[{t}] // {text} {ang} on qubit {t} {cc}"))
    elif name == "SWAP":                       # pairs of targets; an odd one out is dropped (compilable.rs:150-168)
        cq, cc = _controls(controls)
        for k in range(0, len(targets) - 1, 2):
            a, b = targets[k], targets[k + 1]
            out.append((_GATE, f"{cq} swap q[{a}], q[{b}] // SWAP gate between qubits {a} and {b} {cc}"))
    elif name == "Unitary2":
        for t in targets:
            out.append(_unitary_decl(op.matrix, t, controls))
    else:                                      # Matchgate and user operators have no `to_compilable`
        raise CompilerError("UnsupportedOperator", "Operator does not implement Compilable trait")
    return out


def _measurement_instructions(basis, targets: List[int]):
    out = [(_BITREG, len(targets)), (_GROUP, None)]
    for t in targets:
        if basis == MeasurementBasis.Computational:
            out.append((_MEAS, f"measure q[{t}]"))
        elif basis == MeasurementBasis.X:
            out.append((_MEAS, f"xmeasure q[{t}]"))
        elif basis == MeasurementBasis.Y:
            out.append((_MEAS, f"ymeasure q[{t}]"))
        else:                                  # Custom(U): apply U^-1 (conjugate transpose), then measure in Z
            u = basis.matrix
            u_inv = [[complex(u[0][0]).conjugate(), complex(u[1][0]).conjugate()],
                     [complex(u[0][1]).conjugate(), complex(u[1][1]).conjugate()]]
            out.append(_unitary_decl(u_inv, t, []))
            out.append((_MEAS, f"measure q[{t}]"))
    return out


def circuit_to_qasm(circuit, to_dir: Optional[str] = None) -> str:
    """`Circuit::to_qasm` (circuit.rs:244-278): the QASM text; also written to `<to_dir>/circuit.qasm` if given."""
    instructions = []
    for g in circuit.to_concrete_circuit().gates:
        if g.kind == "Operator":
            instructions += _operator_instructions(g.op, list(g.targets), list(g.controls))
        elif g.kind == "Measurement":
            instructions += _measurement_instructions(g.basis, list(g.targets))
        else:                                  # compilable.rs:98-100: `unimplemented!` in the reference
            raise CompilerError("UnsupportedOperator", "Compilation for Pauli time evolution gates is not yet implemented")
    header = ('OPENQASM 3.0;\ninclude "stdgates.inc";\n'
              "def xmeasure(qubit q) -> bit { h q; return measure q; } // Defines x-basis measurement\n"
              "def ymeasure(qubit q) -> bit { h q; s q; return measure q; } // Defines y-basis measurement\n"
              "// Generated by QuantIron's QASM Compiler\n\n"
              f"qubit[{circuit.num_qubits}] q; // Main register\n\n")
    body = ""
    reg, bit, meas = -1, 0, 0
    for kind, val in instructions:
        if kind == _BITREG:
            header += f"bit[{val}] m{meas}; // Measurement result for measurement {meas}\n"
            meas += 1
        elif kind == _GATE:
            if body and not body.endswith("\n"):
                body += "\n"
            body += val.strip() + "\n"
        elif kind == _MEAS:
            body += f"m{reg}[{bit}] = {val};\n"
            bit += 1
        else:                                  # start of a measurement group
            reg, bit = reg + 1, 0
            if body and not body.endswith("\n\n"):
                body += "\n"
            body += f"// Measurement {reg}\n"
    text = header.rstrip()
    if body:
        text += "\n\n"
    text = (text + body.lstrip()).strip()
    if to_dir is not None:
        if not os.path.isdir(to_dir):
            raise CompilerError("IOError", f"Provided path is not a directory: {to_dir}")
        path = os.path.join(to_dir, "circuit.qasm")
        try:
            with open(path, "w") as f:
                f.write(text)
        except OSError as e:
            raise CompilerError("IOError", f"Error creating file '{path}': {e}")
    return text
