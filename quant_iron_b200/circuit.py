"""Gate, Circuit, CircuitBuilder and Subroutine (src/components/gate.rs, src/circuit.rs,
src/subroutine.rs) driving the device engine.

`Circuit::execute` in the reference is a sequential loop with one full-state sweep (plus a clone) per
gate (circuit.rs:160-172).  Here consecutive operator gates are handed to `qi_apply_circuit` as one
run of records; the library validates them, then schedules them into fused register-window passes.
Measurement / PauliString / PauliTimeEvolution gates split the runs and go through their own entry
points.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
from typing import List, Optional, Sequence

from . import _ffi
from .errors import Error
from .measurement import MeasurementBasis
from .operators import (CNOT, SWAP, Hadamard, Identity, Matchgate, Operator, Pauli, PhaseS, PhaseSdag, PhaseShift,
                        PhaseT, PhaseTdag, RotateX, RotateY, RotateZ, Toffoli, Unitary2)
from .parametric import (Parameter, ParametricGate, ParametricMatchgate, ParametricP, ParametricRx, ParametricRy,
                         ParametricRyPhase, ParametricRyPhaseDag, ParametricRz)
from .pauli import PauliString

_lib = _ffi.lib


class Gate:
    """`enum Gate` (gate.rs:13-52).  Parametric gates resolve to Operator gates before they touch
    amplitudes (gate.rs:107-114), so the device path only sees the four kinds below."""

    def __init__(self, kind, **kw):
        self.kind = kind
        self.__dict__.update(kw)

    @staticmethod
    def Operator(op: Operator, targets, controls=()):
        return Gate("Operator", op=op, targets=list(targets), controls=list(controls))

    @staticmethod
    def Measurement(basis, indices):
        return Gate("Measurement", basis=basis, targets=list(indices), controls=[])

    @staticmethod
    def PauliString(ps):
        return Gate("PauliString", pauli_string=ps, targets=ps.get_targets(), controls=[])

    @staticmethod
    def PauliTimeEvolution(ps, time):
        return Gate("PauliTimeEvolution", pauli_string=ps, time=float(time), targets=ps.get_targets(), controls=[])

    @staticmethod
    def Parametric(p_gate, targets, controls=()):
        """gate.rs:34: resolved to concrete operator gates each time it is applied (gate.rs:107-114)."""
        return Gate("Parametric", p_gate=p_gate, targets=list(targets), controls=list(controls))

    new_operator = Operator
    new_measurement = staticmethod(lambda indices, basis: Gate.Measurement(basis, indices))

    def concrete(self) -> "List[Gate]":
        """Circuit::to_concrete_circuit's per-gate rule (circuit.rs:205-217)."""
        if self.kind == "Parametric":
            return self.p_gate.to_concrete_gates(self.targets, self.controls)
        if self.kind == "PauliString":
            return self.pauli_string.to_gates()
        return [self]

    def apply_(self, state, seed: Optional[int] = None):
        """gate.rs:99-122, in place."""
        if self.kind == "Operator":
            return state.apply_(self.op, self.targets, self.controls)
        if self.kind == "Parametric":
            for g in self.p_gate.to_concrete_gates(self.targets, self.controls):
                g.apply_(state)
            return state
        if self.kind == "Measurement":
            state.measure_(self.basis, self.targets, seed=seed)
            return state
        if self.kind == "PauliString":
            self.pauli_string.apply_(state, with_coefficient=False)   # coefficient dropped, gate.rs:115-117
            return state.normalise_()
        if self.pauli_string.coefficient().imag != 0.0:
            raise Error("InvalidPauliStringCoefficient", self.pauli_string.coefficient())
        return self.pauli_string.apply_exp_factor_(state, complex(0.0, -self.time))

    def apply(self, state, seed: Optional[int] = None):
        return self.apply_(state.clone(), seed)

    def get_target_qubits(self):
        return self.targets

    def get_control_qubits(self):
        return self.controls if self.kind in ("Operator", "Parametric") else None

    def __repr__(self):
        if self.kind == "Operator":
            return f"Gate.Operator({self.op!r}, {self.targets}, {self.controls})"
        if self.kind == "Parametric":
            return f"Gate.Parametric({self.p_gate!r}, {self.targets}, {self.controls})"
        return f"Gate.{self.kind}({self.targets})"


def _install_gate_ctors():
    """gate.rs:165-1267 constructor helpers (same names and argument order)."""
    simple = {"h": Hadamard, "x": lambda: Pauli.X, "y": lambda: Pauli.Y, "z": lambda: Pauli.Z, "i": Identity,
              "s": PhaseS, "s_dag": PhaseSdag, "t": PhaseT, "t_dag": PhaseTdag}
    for name, fac in simple.items():
        setattr(Gate, f"{name}_gate", staticmethod(lambda q, _f=fac: Gate.Operator(_f(), [q], [])))
        setattr(Gate, f"{name}_multi_gate", staticmethod(lambda qs, _f=fac: [Gate.Operator(_f(), [q], []) for q in qs]))
        setattr(Gate, f"{name}_controlled_gates",
                staticmethod(lambda ts, cs, _f=fac: [Gate.Operator(_f(), [q], list(cs)) for q in ts]))
    for name, cls in {"p": PhaseShift, "rx": RotateX, "ry": RotateY, "rz": RotateZ}.items():
        setattr(Gate, f"{name}_gate", staticmethod(lambda q, a, _c=cls: Gate.Operator(_c(a), [q], [])))
        setattr(Gate, f"{name}_multi_gate",
                staticmethod(lambda qs, a, _c=cls: [Gate.Operator(_c(a), [q], []) for q in qs]))
        setattr(Gate, f"{name}_controlled_gates",
                staticmethod(lambda ts, cs, a, _c=cls: [Gate.Operator(_c(a), [q], list(cs)) for q in ts]))
    Gate.unitary2_gate = staticmethod(lambda q, u: Gate.Operator(Unitary2.new(u), [q], []))
    Gate.unitary2_multi_gate = staticmethod(lambda qs, u: [Gate.Operator(Unitary2.new(u), [q], []) for q in qs])
    Gate.unitary2_controlled_gates = staticmethod(
        lambda ts, cs, u: [Gate.Operator(Unitary2.new(u), [q], list(cs)) for q in ts])
    Gate.ry_phase_gate = staticmethod(lambda q, th, ph: Gate.Operator(Unitary2.from_ry_phase(th, ph), [q], []))
    Gate.ry_phase_dag_gate = staticmethod(
        lambda q, th, ph: Gate.Operator(Unitary2.from_ry_phase_dagger(th, ph), [q], []))
    Gate.ry_phase_multi_gate = staticmethod(
        lambda qs, th, ph: [Gate.Operator(Unitary2.from_ry_phase(th, ph), [q], []) for q in qs])
    Gate.ry_phase_controlled_gates = staticmethod(
        lambda ts, cs, th, ph: [Gate.Operator(Unitary2.from_ry_phase(th, ph), [q], list(cs)) for q in ts])
    Gate.ry_phase_dag_multi_gate = staticmethod(
        lambda qs, th, ph: [Gate.Operator(Unitary2.from_ry_phase_dagger(th, ph), [q], []) for q in qs])
    Gate.ry_phase_dag_controlled_gates = staticmethod(
        lambda ts, cs, th, ph: [Gate.Operator(Unitary2.from_ry_phase_dagger(th, ph), [q], list(cs)) for q in ts])
    Gate.cnot_gate = staticmethod(lambda target, control: Gate.Operator(CNOT(), [target], [control]))  # gate.rs:1128
    Gate.swap_gate = staticmethod(lambda q1, q2: Gate.Operator(SWAP(), [q1, q2], []))
    Gate.swap_controlled_gate = staticmethod(lambda q1, q2, cs: Gate.Operator(SWAP(), [q1, q2], list(cs)))
    Gate.toffoli_gate = staticmethod(lambda target, controls: Gate.Operator(Toffoli(), [target], list(controls)))
    Gate.pauli_string_gate = staticmethod(lambda ps: Gate.PauliString(ps))
    Gate.pauli_time_evolution_gate = staticmethod(lambda ps, t: Gate.PauliTimeEvolution(ps, t))
    Gate.matchgate = staticmethod(lambda t, th, p1, p2: Gate.Operator(Matchgate(th, p1, p2), [t], []))
    Gate.controlled_matchgate = staticmethod(
        lambda t, cs, th, p1, p2: Gate.Operator(Matchgate(th, p1, p2), [t], list(cs)))


_install_gate_ctors()


class Subroutine:
    """subroutine.rs:13-160."""

    def __init__(self, gates, num_qubits):
        self.gates, self.num_qubits = list(gates), num_qubits

    @staticmethod
    def new(num_qubits):
        return Subroutine([], num_qubits)

    @staticmethod
    def with_gates(gates, num_qubits):
        return Subroutine(gates, num_qubits)

    def get_gates(self):
        return self.gates

    def add_gate(self, gate):
        self.gates.append(gate)

    def add_gates(self, gates):
        self.gates.extend(gates)

    def get_num_qubits(self):
        return self.num_qubits

    @staticmethod
    def qft(qubits: Sequence[int], num_qubits: int) -> "Subroutine":
        """subroutine.rs:90-112: H(q_i), CP(target q_i, control q_{i+k}, pi / 2^k) with the f64
        denominator doubled in a loop (99-104), then the swaps."""
        b = CircuitBuilder(num_qubits)
        n = len(qubits)
        for i in range(n):
            b.h_gate(qubits[i])
            den = 2.0
            for k in range(1, n - i):
                b.cp_gates([qubits[i]], [qubits[i + k]], math.pi / den)
                den *= 2.0
        for i in range(n // 2):
            b.swap_gate(qubits[i], qubits[n - 1 - i])
        return b.build_subroutine()

    @staticmethod
    def iqft(qubits: Sequence[int], num_qubits: int) -> "Subroutine":
        """subroutine.rs:125-160: swaps first; denominator from 2f64.powi(k) then halved (141-153)."""
        b = CircuitBuilder(num_qubits)
        n = len(qubits)
        for i in range(n // 2):
            b.swap_gate(qubits[i], qubits[n - 1 - i])
        for i in reversed(range(n)):
            if n > i + 1:
                k_initial = (n - 1) - i
                den = 2.0 ** k_initial
                for it in range(k_initial):
                    k = k_initial - it
                    b.cp_gates([qubits[i]], [qubits[i + k]], -math.pi / den)
                    if k > 1:
                        den /= 2.0
            b.h_gate(qubits[i])
        return b.build_subroutine()


class Circuit:
    """circuit.rs:27-202."""

    def __init__(self, num_qubits: int):
        self.gates: List[Gate] = []
        self.num_qubits = num_qubits
        self._records = None
        self._records_fp = None

    @staticmethod
    def new(num_qubits):
        return Circuit(num_qubits)

    @staticmethod
    def _validate(gate: Gate, n: int):  # circuit.rs:35-52
        for q in gate.get_target_qubits():
            if q >= n:
                raise Error("InvalidQubitIndex", q, n)
        for q in gate.get_control_qubits() or []:
            if q >= n:
                raise Error("InvalidQubitIndex", q, n)

    @staticmethod
    def with_gates(gates, num_qubits):
        for g in gates:
            Circuit._validate(g, num_qubits)
        c = Circuit(num_qubits)
        c.gates = list(gates)
        return c

    def add_gate(self, gate):
        Circuit._validate(gate, self.num_qubits)
        self.gates.append(gate)
        self._records = None

    def add_gates(self, gates):
        for g in gates:
            Circuit._validate(g, self.num_qubits)
        self.gates.extend(gates)
        self._records = None

    def get_num_qubits(self):
        return self.num_qubits

    def get_gates(self):
        return self.gates

    # ---- lowering to C-ABI records ----
    def _lower(self):
        """Split the gate list into runs: ('ops', qi_gate[count], count, keepalive) |
        ('evol', qi_pauli_term[count], count, keepalive, factors) | ('gate', Gate, index)."""
        # `gates` is a public, mutable list of mutable gates (circuit.rs: `pub gates`): the cached records are only
        # reused while every gate, its qubits and its operator's parameters are what they were when the cache was built
        fp = self._fingerprint()
        if self._records is not None and self._records_fp == fp:
            return self._records
        has_parametric = any(g.kind == "Parametric" for g in self.gates)
        runs = []
        cur: List[Gate] = []

        def close():
            if not cur:
                return
            arr = (_ffi.QiGate * len(cur))()
            keep = []
            for i, g in enumerate(cur):
                rec, k = g.op.record(g.targets, g.controls)
                arr[i] = rec
                keep.append(k)
            runs.append(("ops", arr, len(cur), keep))
            cur.clear()

        evol: List[Gate] = []

        def close_evol():
            # consecutive PauliTimeEvolution gates = one apply_exp_factor sequence (fused passes on the device)
            if not evol:
                return
            arr = (_ffi.QiPauliTerm * len(evol))()
            keep, factors = [], []
            for i, g in enumerate(evol):
                rec, k = g.pauli_string.term()
                arr[i] = rec
                keep.append(k)
                factors += [0.0, -g.time]
            runs.append(("evol", arr, len(evol), (keep, list(evol)), _ffi.dbl_array(factors)))
            evol.clear()

        for index, g in enumerate(self.gates):
            if g.kind == "Operator":
                close_evol()
                cur.append(g)
            elif g.kind == "Parametric":
                # resolved with the parameter values of THIS execution; the concrete gates join the fused run
                close_evol()
                cur.extend(g.p_gate.to_concrete_gates(g.targets, g.controls))
            elif g.kind == "PauliTimeEvolution":
                close()
                evol.append(g)
            else:
                close()
                close_evol()
                runs.append(("gate", g, index))
        close()
        close_evol()
        if not has_parametric:
            self._records, self._records_fp = runs, fp
        return runs

    def _fingerprint(self):
        out = [len(self.gates)]
        for g in self.gates:
            out.append(id(g))
            if g.kind == "Operator":
                out.append((id(g.op), tuple(g.targets), tuple(g.controls), tuple(g.op.params())))
            elif g.kind == "PauliTimeEvolution":
                out.append((id(g.pauli_string), g.time))
        return hash(tuple(out))

    def execute_(self, state, seed: Optional[int] = None):
        """Run the circuit IN PLACE on `state` (what a 33-qubit state needs)."""
        if state.num_qubits != self.num_qubits:
            raise Error("InvalidNumberOfQubits", state.num_qubits)
        for run in self._lower():
            if run[0] == "ops":
                _ffi.check(_lib.qi_apply_circuit(state._h, run[1], run[2]))
            elif run[0] == "evol":
                for g in run[3][1]:                       # gate.rs:116-118 -> pauli_string.rs:281-284
                    if g.pauli_string.coefficient().imag != 0.0:
                        raise Error("InvalidPauliStringCoefficient", g.pauli_string.coefficient())
                _ffi.check(_lib.qi_apply_pauli_exp_sequence(state._h, run[1], run[2], run[4]))
            else:
                # gate k of the circuit draws from the stream seeded `seed + k` (shared-seed contract)
                run[1].apply_(state, None if seed is None else seed + run[2])
        return state

    def execute_host_(self, state, host_in: np.ndarray, host_out: Optional[np.ndarray] = None) -> np.ndarray:
        """Circuit::execute for a HOST-resident state vector (the reference's `State.state_vector`, state.rs:74-81):
        `host_in` -> circuit -> `host_out` (default: a new array; may be `host_in` itself), with `state` as the device
        working buffer.  A circuit that is one run of operator gates goes through qi_execute_host, which overlaps the
        two PCIe copies with the circuit (csrc/host_pipeline.cu); anything else is upload, execute_, download."""
        if state.num_qubits != self.num_qubits:
            raise Error("InvalidNumberOfQubits", state.num_qubits)
        hin = np.ascontiguousarray(host_in, dtype=np.complex128)
        if host_out is None:
            host_out = np.empty_like(hin)
        if host_out.dtype != np.complex128 or not host_out.flags.c_contiguous or host_out.shape != hin.shape:
            raise ValueError("host_out must be a contiguous complex128 array of the state's length")
        runs = self._lower()
        if len(runs) == 1 and runs[0][0] == "ops":
            _ffi.check(_lib.qi_execute_host(state._h, runs[0][1], runs[0][2], hin.ctypes.data_as(C.c_void_p),
                                            host_out.ctypes.data_as(C.c_void_p), hin.shape[0]))
            return host_out
        state.upload_(hin)
        self.execute_(state)
        return state.to_host(host_out)

    def execute(self, initial_state, seed: Optional[int] = None):  # circuit.rs:160-172
        if initial_state.num_qubits != self.num_qubits:
            raise Error("InvalidNumberOfQubits", initial_state.num_qubits)
        return self.execute_(initial_state.clone(), seed)

    def to_concrete_circuit(self) -> "Circuit":  # circuit.rs:204-221
        c = Circuit(self.num_qubits)
        c.gates = [cg for g in self.gates for cg in g.concrete()]
        return c

    def to_qasm(self, to_dir=None) -> str:  # circuit.rs:244-278 (host-only string emission, qasm.py)
        from .qasm import circuit_to_qasm
        return circuit_to_qasm(self, to_dir)

    def trace_execution(self, initial_state):  # circuit.rs:188-202
        if initial_state.num_qubits != self.num_qubits:
            raise Error("InvalidNumberOfQubits", initial_state.num_qubits)
        cur = initial_state.clone()
        out = [cur.clone()]
        for g in self.gates:
            g.apply_(cur)
            out.append(cur.clone())
        return out


class CircuitBuilder:
    """circuit.rs:288-1742 (operator, Pauli and measurement adders)."""

    def __init__(self, num_qubits: int):
        self.gates: List[Gate] = []
        self.num_qubits = num_qubits

    @staticmethod
    def new(num_qubits):
        return CircuitBuilder(num_qubits)

    def add_gate(self, gate):
        self.gates.append(gate)
        return self

    def add_gates(self, gates):
        self.gates.extend(gates)
        return self

    def build(self) -> Circuit:  # circuit.rs:340-343
        return Circuit.with_gates(self.gates, self.num_qubits)

    def build_final(self) -> Circuit:
        c = Circuit.with_gates(self.gates, self.num_qubits)
        self.gates = []
        return c

    def build_subroutine(self) -> Subroutine:
        s = Subroutine(self.gates, self.num_qubits)
        self.gates = []
        return s

    def add_subroutine(self, sub: Subroutine):  # circuit.rs:373-376
        self.gates.extend(sub.gates)
        return self

    def _each(self, op, targets, controls=()):
        for q in targets:
            self.gates.append(Gate.Operator(op, [q], controls))
        return self

    def cnot_gate(self, target_qubit, control_qubit):  # circuit.rs:1071: TARGET first
        return self.add_gate(Gate.Operator(CNOT(), [target_qubit], [control_qubit]))

    def swap_gate(self, q1, q2):
        return self.add_gate(Gate.Operator(SWAP(), [q1, q2], []))

    def cswap_gate(self, t1, t2, controls):
        return self.add_gate(Gate.Operator(SWAP(), [t1, t2], controls))

    def toffoli_gate(self, control1, control2, target):  # circuit.rs:1118-1123
        return self.add_gate(Gate.Operator(Toffoli(), [target], [control1, control2]))

    def pauli_string_gate(self, ps: PauliString):
        return self.add_gate(Gate.PauliString(ps))

    def pauli_time_evolution_gate(self, ps: PauliString, time: float):
        return self.add_gate(Gate.PauliTimeEvolution(ps, time))

    def matchgate(self, target, theta, phi1, phi2):
        return self.add_gate(Gate.Operator(Matchgate(theta, phi1, phi2), [target], []))

    def cmatchgate(self, target, controls, theta, phi1, phi2):  # circuit.rs:1194-1200: controls SECOND (State::cmatchgate has them last)
        return self.add_gate(Gate.Operator(Matchgate(theta, phi1, phi2), [target], list(controls)))

    def add_operator_gate(self, op, targets, controls=()):  # circuit.rs:1215-1224
        return self.add_gate(Gate.Operator(op, targets, controls))

    def unitary_gate(self, qubit, unitary):
        return self.add_gate(Gate.Operator(Unitary2.new(unitary), [qubit], []))

    def unitary_gates(self, qubits, unitary):
        return self._each(Unitary2.new(unitary), qubits)

    def cunitary_gates(self, targets, controls, unitary):
        return self._each(Unitary2.new(unitary), targets, controls)

    def ry_phase_gate(self, qubit, theta, phi):
        return self.add_gate(Gate.Operator(Unitary2.from_ry_phase(theta, phi), [qubit], []))

    def ry_phase_gates(self, qubits, theta, phi):
        return self._each(Unitary2.from_ry_phase(theta, phi), qubits)

    def cry_phase_gates(self, targets, controls, theta, phi):
        return self._each(Unitary2.from_ry_phase(theta, phi), targets, controls)

    def ry_phase_dag_gate(self, qubit, theta, phi):
        return self.add_gate(Gate.Operator(Unitary2.from_ry_phase_dagger(theta, phi), [qubit], []))

    def ry_phase_dag_gates(self, qubits, theta, phi):
        return self._each(Unitary2.from_ry_phase_dagger(theta, phi), qubits)

    def cry_phase_dag_gates(self, targets, controls, theta, phi):
        return self._each(Unitary2.from_ry_phase_dagger(theta, phi), targets, controls)

    def measure_gate(self, basis: MeasurementBasis, qubits):
        return self.add_gate(Gate.Measurement(basis, qubits))

    # ---- parametric gates (circuit.rs:1226-1742) ----
    def _parametric_each(self, cls, targets, controls, parameters):
        targets, parameters = list(targets), list(parameters)
        if len(targets) != len(parameters):
            raise Error("MismatchedNumberOfParameters", len(targets), len(parameters))
        for t, prm in zip(targets, parameters):
            self.add_gate(Gate.Parametric(cls(prm), [t], list(controls)))
        return self

    def parametric_matchgate(self, target_index, parameter):
        return self.add_gate(Gate.Parametric(ParametricMatchgate(parameter), [target_index], []))

    def parametric_cmatchgate(self, target_index, control_indices, parameter):
        return self.add_gate(Gate.Parametric(ParametricMatchgate(parameter), [target_index], list(control_indices)))


def _install_builder_methods():
    simple = {"h": Hadamard, "x": lambda: Pauli.X, "y": lambda: Pauli.Y, "z": lambda: Pauli.Z,
              "s": PhaseS, "t": PhaseT, "sdag": PhaseSdag, "tdag": PhaseTdag}
    for name, fac in simple.items():
        setattr(CircuitBuilder, f"{name}_gate", lambda self, q, _f=fac: self._each(_f(), [q]))
        setattr(CircuitBuilder, f"{name}_gates", lambda self, qs, _f=fac: self._each(_f(), qs))
        setattr(CircuitBuilder, f"c{name}_gates", lambda self, t, c, _f=fac: self._each(_f(), t, c))
    CircuitBuilder.id_gate = lambda self, q: self._each(Identity(), [q])
    CircuitBuilder.id_gates = lambda self, qs: self._each(Identity(), qs)
    CircuitBuilder.ci_gates = lambda self, t, c: self._each(Identity(), t, c)
    for name, cls in {"p": PhaseShift, "rx": RotateX, "ry": RotateY, "rz": RotateZ}.items():
        setattr(CircuitBuilder, f"{name}_gate", lambda self, q, a, _c=cls: self._each(_c(a), [q]))
        setattr(CircuitBuilder, f"{name}_gates", lambda self, qs, a, _c=cls: self._each(_c(a), qs))
        setattr(CircuitBuilder, f"c{name}_gates", lambda self, t, c, a, _c=cls: self._each(_c(a), t, c))
    # parametric adders (circuit.rs:1226-1742): <name>_gate(target, parameter), <name>_gates(targets, parameters),
    # c<name>_gates(targets, controls, parameters) -- one Parameter per target
    for name, cls in {"ry_phase": ParametricRyPhase, "ry_phase_dag": ParametricRyPhaseDag, "rx": ParametricRx,
                      "ry": ParametricRy, "rz": ParametricRz, "p": ParametricP}.items():
        setattr(CircuitBuilder, f"parametric_{name}_gate",
                lambda self, t, prm, _c=cls: self.add_gate(Gate.Parametric(_c(prm), [t], [])))
        setattr(CircuitBuilder, f"parametric_{name}_gates",
                lambda self, ts, prms, _c=cls: self._parametric_each(_c, ts, [], prms))
        setattr(CircuitBuilder, f"parametric_c{name}_gates",
                lambda self, ts, cs, prms, _c=cls: self._parametric_each(_c, ts, cs, prms))


_install_builder_methods()
