"""quant_iron_b200 -- B200-native state-vector engine behind quant-iron's API.

Host-side mirror of the reference's public surface for the state-vector hot path (State gate
methods, the Operator trait, Gate/Circuit/CircuitBuilder/Subroutine::qft, PauliString/SumOp apply,
exp and expectation, measure/measure_n, Trotter, Heisenberg) over the C ABI of libqiron_b200.so
(include/qiron_b200.h): hand-written sm_100a CUDA kernels on a device-resident Complex<f64> buffer.
No OpenCL, no multi-backend dispatch, no CPU fallback.
"""
from . import _ffi
from .algorithms import (TrotterOrder, first_order_trotter_step, second_order_trotter_step, trotter_evolve_state,
                         trotter_evolve_state_)
from .circuit import Circuit, CircuitBuilder, Gate, Subroutine
from .errors import CompilerError, Error
from .measurement import MeasurementBasis, MeasurementResult
from .models import heisenberg_1d, heisenberg_2d, ising_1d, ising_1d_uniform, ising_2d, ising_2d_uniform
from .operators import (CNOT, SWAP, Hadamard, Identity, Matchgate, Operator, Pauli, PhaseS, PhaseSdag, PhaseShift,
                        PhaseT, PhaseTdag, RotateX, RotateY, RotateZ, Toffoli, Unitary2)
from .parametric import (Parameter, ParametricGate, ParametricMatchgate, ParametricP, ParametricRx, ParametricRy,
                         ParametricRyPhase, ParametricRyPhaseDag, ParametricRz)
from .pauli import PauliString, SumOp
from .state import State
from .chain import ChainableState, chain
from . import workloads
from . import engine
from . import macros
from .macros import circuit as circuit_macro

__all__ = [
    "State", "ChainableState", "chain", "Operator", "Hadamard", "Pauli", "CNOT", "SWAP", "Toffoli", "Identity", "PhaseS", "PhaseT",
    "PhaseSdag", "PhaseTdag", "PhaseShift", "RotateX", "RotateY", "RotateZ", "Unitary2", "Matchgate",
    "Gate", "Circuit", "CircuitBuilder", "Subroutine", "PauliString", "SumOp", "MeasurementBasis",
    "MeasurementResult", "TrotterOrder", "first_order_trotter_step", "second_order_trotter_step",
    "trotter_evolve_state", "trotter_evolve_state_", "Parameter", "ParametricGate", "ParametricMatchgate", "ParametricP", "ParametricRx", "ParametricRy", "ParametricRyPhase",
    "ParametricRyPhaseDag", "ParametricRz", "heisenberg_1d", "heisenberg_2d", "ising_1d", "ising_1d_uniform", "ising_2d", "ising_2d_uniform", "Error", "CompilerError", "workloads", "engine", "macros", "circuit_macro",
]
