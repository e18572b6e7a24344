"""Sharded state over 2/4/8 GPUs of one node: one process per GPU (torchrun), torch.distributed
only for the plumbing (exchanging the CUDA IPC handles of every rank's amplitude buffer; gathering
small results in tests).  All amplitude traffic between GPUs is done by the library's own kernels over
peer-mapped memory (csrc/shard.cu): there is no NCCL call on the data path.

    import torch.distributed as dist
    dist.init_process_group("nccl")            # or gloo: only small host objects travel here
    st = sharded.new_zero(34, dist)            # 34 logical qubits, top log2(world) are global
    circuit.execute_(st)                       # same Circuit / gate API as a single-GPU State
"""
from __future__ import annotations

import ctypes as C
from typing import List

import numpy as np

from . import _ffi
from .state import State

_lib = _ffi.lib
HANDLE_BYTES = 2 * _ffi.IPC_HANDLE_BYTES


def exchange_handles(local: bytes, dist) -> List[bytes]:
    """All-gather one fixed-size byte string per rank (works with gloo and nccl process groups)."""
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, local)
    return out


def _attach(handle, dist) -> State:
    buf = (C.c_uint8 * HANDLE_BYTES)()
    _ffi.check(_lib.qi_shard_export(handle, buf))
    handles = exchange_handles(bytes(buf), dist)
    flat = b"".join(handles)
    arr = (C.c_uint8 * len(flat)).from_buffer_copy(flat)
    _ffi.check(_lib.qi_shard_attach(handle, arr))
    st = State(_handle=handle)
    dist.barrier()
    return st


def _new(fn, dist, *args) -> State:
    h = C.c_void_p()
    _ffi.check(fn(*args, dist.get_rank(), dist.get_world_size(), C.byref(h)))
    return _attach(h, dist)


def new_zero(num_qubits: int, dist) -> State:
    return _new(_lib.qi_shard_new_zero, dist, num_qubits)


def new_plus(num_qubits: int, dist) -> State:
    return _new(_lib.qi_shard_new_plus, dist, num_qubits)


def new_basis_n(num_qubits: int, n: int, dist) -> State:
    return _new(_lib.qi_shard_new_basis_n, dist, num_qubits, n)


def layout(state: State):
    """(phys[logical qubit] -> physical bit, n_local)"""
    phys = (C.c_uint8 * 64)()
    nl = C.c_uint32()
    _ffi.check(_lib.qi_state_layout(state._h, phys, C.byref(nl)))
    return [int(phys[q]) for q in range(state.num_qubits)], int(nl.value)


def unpermute(physical: np.ndarray, phys: List[int]) -> np.ndarray:
    """Reorder amplitudes stored in physical bit order into logical order: out[i] = physical[p(i)],
    p(i) = sum_q bit_q(i) << phys[q]."""
    n = len(phys)
    idx = np.arange(1 << n, dtype=np.uint64)
    pidx = np.zeros_like(idx)
    for q in range(n):
        pidx |= ((idx >> np.uint64(q)) & np.uint64(1)) << np.uint64(phys[q])
    return physical[pidx]


def shard_to_host(state: State, out: np.ndarray = None) -> np.ndarray:
    """This rank's shard as stored (PHYSICAL bit order; `layout` gives the logical -> physical qubit map)."""
    if out is None:
        out = np.empty(len(state), dtype=np.complex128)
    _ffi.check(_lib.qi_shard_to_host(state._h, out.ctypes.data_as(C.c_void_p), out.shape[0]))
    return out


def gather_state_vector(state: State, dist) -> np.ndarray:
    """Every rank's shard, concatenated in rank order and put back into logical qubit order
    (test helper: only for states that fit host memory)."""
    local = shard_to_host(state)
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, local)
    phys, _ = layout(state)
    return unpermute(np.concatenate(parts), phys)


def comm_stats(state: State) -> dict:
    a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
    _ffi.check(_lib.qi_shard_comm_stats(state._h, C.byref(a), C.byref(b), C.byref(c)))
    return {"bytes_sent": a.value, "bytes_received": b.value, "exchanges": c.value}


def plan(num_qubits: int, world: int, circuit) -> dict:
    """Host-only: how many global<->local exchanges `circuit` costs on `world` ranks (no GPU needed)."""
    records = []
    for g in circuit.gates:
        if g.kind != "Operator":
            raise ValueError("plan() takes circuits of operator gates")
        records.append(g.op.record(g.targets, g.controls))
    arr = (_ffi.QiGate * max(1, len(records)))()
    for i, (rec, _keep) in enumerate(records):
        arr[i] = rec
    ex, free = C.c_uint64(), C.c_uint64()
    phys = (C.c_uint8 * 64)()
    _ffi.check(_lib.qi_shard_plan(num_qubits, world, arr, len(records), C.byref(ex), C.byref(free), phys))
    return {"exchanges": int(ex.value), "comm_free_global_gates": int(free.value),
            "final_layout": [int(phys[q]) for q in range(num_qubits)]}


def plan_pauli(total_qubits: int, world: int, hamiltonian, repeats: int = 1) -> dict:
    """Host-only: exchanges / stages the engine needs for `repeats` repetitions of the term sequence of a SumOp
    (e.g. first-order Trotter steps) on `world` ranks (qi_shard_plan_pauli; no device access)."""
    arr, n, keep = hamiltonian.term_array()
    ex, st = C.c_uint64(0), C.c_uint64(0)
    _ffi.check(_ffi.lib.qi_shard_plan_pauli(total_qubits, world, arr, n, repeats, C.byref(ex), C.byref(st)))
    return {"exchanges": int(ex.value), "stages": int(st.value)}
