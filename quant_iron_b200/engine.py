"""Engine-level controls: device binding, options, kernel accounting, stream stopwatch."""
from __future__ import annotations

import ctypes as C

from . import _ffi

_lib = _ffi.lib


def init(device: int = -1):
    _ffi.check(_lib.qi_init(device))


def synchronize():
    _ffi.check(_lib.qi_synchronize())


def version() -> str:
    return _lib.qi_version().decode()


def device_info() -> dict:
    name = C.create_string_buffer(128)
    sm = C.c_int()
    tot, free = C.c_uint64(), C.c_uint64()
    _ffi.check(_lib.qi_device_info(name, 128, C.byref(sm), C.byref(tot), C.byref(free)))
    return {"name": name.value.decode(), "sm_count": sm.value, "total_mem": tot.value, "free_mem": free.value}


def set_option(name: str, value: int):
    _ffi.check(_lib.qi_set_option(name.encode(), int(value)))


def jit_drain():
    """Wait until every queued tile module is assembled (csrc/tile_jit.cuh)."""
    _ffi.check(_lib.qi_jit_drain())


def jit_stats() -> dict:
    m, f, p, ms, wi = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_double(), C.c_double()
    _ffi.check(_lib.qi_jit_stats(C.byref(m), C.byref(f), C.byref(p), C.byref(ms), C.byref(wi)))
    return {"modules": m.value, "failed": f.value, "pending": p.value, "assemble_ms": ms.value, "fp64_warp_instr": wi.value}


def stats_reset():
    _ffi.check(_lib.qi_stats_reset())


def stats() -> dict:
    arr = (_ffi.QiKernelStat * 32)()
    n = C.c_int()
    _ffi.check(_lib.qi_stats_get(arr, 32, C.byref(n)))
    return {arr[i].name.decode(): {"launches": int(arr[i].launches), "total_ms": float(arr[i].total_ms),
                                   "algorithmic_bytes": float(arr[i].algorithmic_bytes)} for i in range(n.value)}


def timer_start():
    _ffi.check(_lib.qi_timer_start())


def timer_stop() -> float:
    ms = C.c_float()
    _ffi.check(_lib.qi_timer_stop(C.byref(ms)))
    return float(ms.value)


def uniform(seed: int, k: int = 0) -> float:
    return float(_lib.qi_uniform(C.c_uint64(seed & (2**64 - 1)), C.c_uint64(k)))
