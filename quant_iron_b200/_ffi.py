"""ctypes binding of libqiron_b200.so (the C ABI in include/qiron_b200.h).

There is deliberately no fallback: if the shared library is missing, importing the package fails
loudly, and every compute entry returns QI_ERR_CUDA when no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os

from .errors import Error

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QIRON_B200_LIB") or os.path.join(_PKG, "lib", "libqiron_b200.so")  # env override: kernel experiments

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(or `python quant_iron_b200/_build.py`). quant_iron_b200 has no CPU or PyTorch fallback.")

lib = C.CDLL(LIB_PATH)

u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
u8p = C.POINTER(C.c_uint8)
dp = C.POINTER(C.c_double)
state_p = C.c_void_p


class QiGate(C.Structure):
    _fields_ = [("kind", C.c_int32), ("num_targets", C.c_uint32), ("targets", C.c_uint32 * 2),
                ("num_controls", C.c_uint32), ("controls", u32p), ("params", C.c_double * 8)]


class QiPauliTerm(C.Structure):
    _fields_ = [("num_ops", C.c_uint32), ("qubits", u32p), ("paulis", u8p), ("coefficient", C.c_double * 2)]


class QiKernelStat(C.Structure):
    _fields_ = [("name", C.c_char * 32), ("launches", C.c_uint64), ("total_ms", C.c_double),
                ("algorithmic_bytes", C.c_double)]


def _sig(name, argtypes, restype=C.c_int):
    f = getattr(lib, name)
    f.argtypes = argtypes
    f.restype = restype
    return f


_sig("qi_last_error", [u64p, C.c_char_p, C.c_size_t], None)
_sig("qi_version", [], C.c_char_p)
_sig("qi_init", [C.c_int])
_sig("qi_synchronize", [])
_sig("qi_jit_drain", [])
_sig("qi_jit_stats", [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_double), C.POINTER(C.c_double)])
_sig("qi_device_info", [C.c_char_p, C.c_size_t, C.POINTER(C.c_int), u64p, u64p])
_sig("qi_set_option", [C.c_char_p, C.c_int64])
_sig("qi_stats_reset", [])
_sig("qi_stats_get", [C.POINTER(QiKernelStat), C.c_int, C.POINTER(C.c_int)])
_sig("qi_timer_start", [])
_sig("qi_timer_stop", [C.POINTER(C.c_float)])
for _n in ("zero", "plus", "minus", "ghz"):
    _sig(f"qi_state_new_{_n}", [C.c_uint32, C.POINTER(state_p)])
_sig("qi_state_new_basis_n", [C.c_uint32, C.c_uint64, C.POINTER(state_p)])
_sig("qi_state_from_host", [C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.POINTER(state_p)])
_sig("qi_state_to_host", [state_p, C.c_void_p, C.c_uint64])
_sig("qi_shard_to_host", [state_p, C.c_void_p, C.c_uint64])
_sig("qi_state_upload", [state_p, C.c_void_p, C.c_uint64])
_sig("qi_state_clone", [state_p, C.POINTER(state_p)])
_sig("qi_state_free", [state_p], None)
_sig("qi_state_num_qubits", [state_p], C.c_uint32)
_sig("qi_state_len", [state_p], C.c_uint64)
_sig("qi_state_amplitude", [state_p, C.c_uint64, dp])
_sig("qi_state_init_random", [state_p, C.c_uint64])
_sig("qi_state_device_ptr", [state_p], C.c_void_p)
_sig("qi_inner_product", [state_p, state_p, dp])
_sig("qi_norm_sqr", [state_p, dp])
_sig("qi_normalise", [state_p])
_sig("qi_scale", [state_p, dp])
_sig("qi_add", [state_p, state_p])
_sig("qi_sub", [state_p, state_p])
_sig("qi_conj", [state_p])
_sig("qi_tensor_product", [state_p, state_p, C.POINTER(state_p)])
_sig("qi_apply_gate", [state_p, C.POINTER(QiGate)])
_sig("qi_apply_circuit", [state_p, C.POINTER(QiGate), C.c_uint64])
_sig("qi_execute_host", [state_p, C.POINTER(QiGate), C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64])
_sig("qi_host_pipeline_plan", [C.c_uint32, C.POINTER(QiGate), C.c_uint64, C.c_int, u64p, u64p, u64p, u64p])
_sig("qi_unitary2_check", [dp])
_sig("qi_apply_pauli_string", [state_p, C.POINTER(QiPauliTerm), C.c_int])
_sig("qi_apply_pauli_exp", [state_p, C.POINTER(QiPauliTerm), dp])
_sig("qi_expect_pauli_sum", [state_p, C.POINTER(QiPauliTerm), C.c_uint64, dp])
_sig("qi_apply_pauli_sum", [state_p, C.POINTER(QiPauliTerm), C.c_uint64, C.POINTER(state_p)])
_sig("qi_trotter_evolve", [state_p, C.POINTER(QiPauliTerm), C.c_uint64, C.c_double, C.c_uint64, C.c_int])
_sig("qi_apply_pauli_exp_sequence", [state_p, C.POINTER(QiPauliTerm), C.c_uint64, dp])
_sig("qi_debug_pauli_schedule", [C.c_uint32, C.POINTER(QiPauliTerm), C.c_uint64, C.c_uint64, C.POINTER(C.c_int32), C.c_uint64, u64p])
_sig("qi_probabilities", [state_p, u32p, C.c_uint32, dp])
_sig("qi_sample", [state_p, u32p, C.c_uint32, C.c_uint64, C.c_uint64, u64p])
_sig("qi_collapse", [state_p, u32p, C.c_uint32, C.c_uint64])
_sig("qi_measure", [state_p, C.c_int, dp, u32p, C.c_uint32, C.c_uint64, C.c_uint64, u8p, u64p])
_sig("qi_uniform", [C.c_uint64, C.c_uint64], C.c_double)
_sig("qi_shard_new_zero", [C.c_uint32, C.c_int, C.c_int, C.POINTER(state_p)])
_sig("qi_shard_new_plus", [C.c_uint32, C.c_int, C.c_int, C.POINTER(state_p)])
_sig("qi_shard_new_basis_n", [C.c_uint32, C.c_uint64, C.c_int, C.c_int, C.POINTER(state_p)])
_sig("qi_shard_export", [state_p, u8p])
_sig("qi_shard_attach", [state_p, u8p])
_sig("qi_shard_rank", [state_p], C.c_int)
_sig("qi_shard_world", [state_p], C.c_int)
_sig("qi_shard_comm_stats", [state_p, u64p, u64p, u64p])
_sig("qi_debug_schedule", [C.c_uint32, C.POINTER(QiGate), C.c_uint64, C.c_int, C.POINTER(C.c_int32), C.c_uint64, u64p])
_sig("qi_debug_lower", [C.c_uint32, C.c_int, C.c_int, u8p, C.POINTER(QiGate), C.c_uint64, C.c_int, u8p, C.c_uint64, u64p])
_sig("qi_debug_shard_stages", [C.c_uint32, C.c_int, C.POINTER(QiGate), C.c_uint64, u64p, C.c_uint64, u64p])
_sig("qi_debug_pauli_lower", [C.c_uint32, C.c_int, C.c_int, u8p, C.POINTER(QiPauliTerm), C.c_uint64, dp, u8p, C.c_uint64, u64p])
_sig("qi_debug_shard_pauli_stages", [C.c_uint32, C.c_int, C.POINTER(QiPauliTerm), C.c_uint64, u64p, C.c_uint64, u64p])
_sig("qi_debug_expect_lower", [C.c_uint32, C.POINTER(QiPauliTerm), C.c_uint64, u8p, C.c_uint64, u64p])
_sig("qi_state_layout", [state_p, u8p, C.POINTER(C.c_uint32)])
_sig("qi_shard_plan_pauli", [C.c_uint32, C.c_int, C.POINTER(QiPauliTerm), C.c_uint64, C.c_uint64, u64p, u64p])
_sig("qi_shard_plan", [C.c_uint32, C.c_int, C.POINTER(QiGate), C.c_uint64, u64p, u64p, u8p])

IPC_HANDLE_BYTES = 64

_VARIANTS = {
    1: ("InvalidNumberOfMeasurements", 1), 2: ("OverlappingControlAndTargetQubits", 2),
    3: ("InvalidNumberOfQubits", 1), 4: ("InvalidQubitIndex", 2), 5: ("StateVectorNotNormalised", 0),
    6: ("NonUnitaryMatrix", 0), 7: ("InvalidNumberOfInputs", 2), 8: ("MismatchedNumberOfParameters", 2),
    9: ("UnknownError", 0), 10: ("CudaError", 1), 11: ("GpuContextLockError", 0), 12: ("CircuitMacroError", 0),
    13: ("InvalidInputValue", 1), 14: ("ZeroNorm", 0), 15: ("InvalidPauliStringCoefficient", 2),
    16: ("InvalidArgument", 0), 17: ("PeerError", 0),
}


def check(status: int):
    """Translate a qi_status into the reference's Error variant (errors.rs:3-97)."""
    if status == 0:
        return
    payload = (C.c_uint64 * 2)()
    msg = C.create_string_buffer(256)
    lib.qi_last_error(payload, msg, 256)
    name, npay = _VARIANTS.get(status, ("UnknownError", 0))
    err = Error(name, *[int(payload[i]) for i in range(npay)])
    err.message = msg.value.decode(errors="replace")
    err.args = (f"{name}{err.payload}: {err.message}",)
    raise err


def u32_array(xs):
    xs = list(xs)
    return (C.c_uint32 * max(1, len(xs)))(*xs)


def dbl_array(xs):
    xs = list(xs)
    return (C.c_double * max(1, len(xs)))(*xs)
