"""The sharded executor's host half on the CPU: `qi_debug_shard_stages` reports the stages
apply_circuit_sharded runs on `world` ranks (gates per stage, the qubit map each stage runs under, the global<->local
exchange that follows), `qi_debug_lower` the device programs every rank would launch for a stage; this test replays
them with the numpy interpreter (tests/window_interp.py) -- every rank on its slice of one big vector, the exchange as
the bit permutation it is -- and compares the un-permuted result with the oracle.  Covers rank-bit controls and phases,
staging / deferral, exchange planning, eviction and lazy SWAP relabelling without any GPU; what stays GPU-only is the
peer-memory exchange kernel and the device barrier (tests/sharded_worker.py on 2-8 GPUs)."""
import ctypes as C

import numpy as np
import pytest

import window_interp as wi
from conftest import AMP_TOL, vec
from test_host_pipeline import _fuzz_builders


def _stages(circuit, n, world):
    from quant_iron_b200 import _ffi
    runs = circuit._lower()
    assert len(runs) == 1 and runs[0][0] == "ops"
    used = C.c_uint64()
    cap = 1 << 16
    while True:
        buf = (C.c_uint64 * cap)()
        st = _ffi.lib.qi_debug_shard_stages(n, world, runs[0][1], runs[0][2], buf, cap, C.byref(used))
        if st == 0:
            break
        if used.value > cap:
            cap = used.value
            continue
        _ffi.check(st)
    rec = [int(x) for x in buf[:used.value]]
    pos = 1
    stages = []
    for _ in range(rec[0]):
        phys = rec[pos:pos + 64]; pos += 64
        nt = rec[pos]; pos += 1
        take = rec[pos:pos + nt]; pos += nt
        nex = rec[pos]; pos += 1
        G = rec[pos:pos + nex]; pos += nex
        L = rec[pos:pos + nex]; pos += nex
        stages.append((phys, take, G, L))
    final_phys = rec[pos:pos + 64]
    assert pos + 64 == len(rec)
    return stages, final_phys


def _emulate(cg, n, world, start):
    import quant_iron_b200 as gpu
    stages, final_phys = _stages(cg, n, world)
    p = world.bit_length() - 1
    nl = n - p
    V = np.array(start, dtype=np.complex128)
    idx = np.arange(1 << n, dtype=np.uint64)
    exchanges = 0
    for phys, take, G, L in stages:
        if take:
            c = gpu.Circuit.with_gates([cg.gates[i] for i in take], n)
            for r in range(world):
                out_phys, _ = wi.execute(wi.lower(c, n, rank=r, world=world, phys=phys), V[r << nl:(r + 1) << nl], nl)
                assert out_phys[:n] == phys[:n]
        if G:
            assert all(g >= nl for g in G) and all(0 <= l < nl for l in L) and len(set(G)) == len(G) and len(set(L)) == len(L)
            src = idx.copy()
            for g, l in zip(G, L):
                bg, bl = (src >> np.uint64(g)) & np.uint64(1), (src >> np.uint64(l)) & np.uint64(1)
                diff = bg ^ bl
                src ^= (diff << np.uint64(g)) | (diff << np.uint64(l))
            V = V[src]
            exchanges += 1
    return wi.to_logical(V, n, final_phys), exchanges, len(stages)


@pytest.mark.parametrize("n,world,seed", [(11, 2, 1), (12, 4, 2), (12, 8, 3), (13, 8, 4), (11, 8, 5), (12, 2, 6), (13, 4, 7)])
def test_sharded_fuzz_replayed_on_cpu(ref, n, world, seed):
    import quant_iron_b200 as gpu
    cg, cr = _fuzz_builders([gpu, ref], n, seed, count=140, lazy_swaps=(seed % 2 == 1))
    start = ref.random_state(n, 800 + seed)
    got, exchanges, nstages = _emulate(cg, n, world, start.state_vector)
    assert float(np.max(np.abs(got - vec(cr.execute(start))))) <= AMP_TOL
    assert exchanges >= 1 and nstages >= exchanges


@pytest.mark.parametrize("n,world", [(12, 2), (13, 4), (14, 8)])
def test_sharded_layered_circuit_replayed_on_cpu(ref, n, world):
    """The benchmark generator on a sharded state: the light-cone staging needs far fewer exchanges than layers."""
    import quant_iron_b200 as gpu
    from quant_iron_b200 import workloads as w
    depth = 12
    specs = w.random_layered_circuit(n, depth)
    start = ref.random_state(n, 2)
    got, exchanges, _ = _emulate(w.build_circuit(gpu, n, specs), n, world, start.state_vector)
    want = vec(w.build_circuit(ref, n, specs).execute(start))
    assert float(np.max(np.abs(got - want))) <= AMP_TOL
    assert 1 <= exchanges < depth


@pytest.mark.parametrize("n,world", [(12, 8), (13, 4)])
def test_sharded_qft_replayed_on_cpu(ref, n, world):
    """QFT on a sharded state: controlled phases on rank-bit qubits are communication-free, the Hadamards on the global
    qubits come in with ONE exchange, the final swaps are relabelled; QFT|+..+> = |0..0>."""
    import quant_iron_b200 as gpu
    from quant_iron_b200 import workloads as w
    c = w.build_circuit(gpu, n, w.qft_specs(n))
    plus = np.full(1 << n, 1.0 / np.sqrt(float(1 << n)), dtype=np.complex128)
    got, exchanges, _ = _emulate(c, n, world, plus)
    assert abs(got[0] - 1.0) <= AMP_TOL and float(np.max(np.abs(got[1:]))) <= AMP_TOL
    assert exchanges == 1
    start = ref.random_state(n, 6)
    got, _, _ = _emulate(c, n, world, start.state_vector)
    assert float(np.max(np.abs(got - vec(w.build_circuit(ref, n, w.qft_specs(n)).execute(start))))) <= AMP_TOL


def _pauli_stages(strings, n, world):
    from quant_iron_b200 import _ffi
    arr = (_ffi.QiPauliTerm * len(strings))()
    keep = []
    for i, ps in enumerate(strings):
        rec, k = ps.term()
        arr[i] = rec
        keep.append(k)
    used = C.c_uint64()
    cap = 1 << 16
    while True:
        buf = (C.c_uint64 * cap)()
        st = _ffi.lib.qi_debug_shard_pauli_stages(n, world, arr, len(strings), buf, cap, C.byref(used))
        if st == 0:
            break
        if used.value > cap:
            cap = used.value
            continue
        _ffi.check(st)
    rec = [int(x) for x in buf[:used.value]]
    pos = 1
    stages = []
    for _ in range(rec[0]):
        phys = rec[pos:pos + 64]; pos += 64
        nt = rec[pos]; pos += 1
        take = rec[pos:pos + nt]; pos += nt
        nex = rec[pos]; pos += 1
        G = rec[pos:pos + nex]; pos += nex
        L = rec[pos:pos + nex]; pos += nex
        stages.append((phys, take, G, L))
    final_phys = rec[pos:pos + 64]
    assert pos + 64 == len(rec)
    return stages, final_phys


def _emulate_pauli(strings, factors, n, world, start):
    stages, final_phys = _pauli_stages(strings, n, world)
    nl = n - (world.bit_length() - 1)
    V = np.array(start, dtype=np.complex128)
    idx = np.arange(1 << n, dtype=np.uint64)
    exchanges = 0
    for phys, take, G, L in stages:
        if take:
            sub, fac = [strings[i] for i in take], [factors[i] for i in take]
            for r in range(world):
                view = V[r << nl:(r + 1) << nl]
                wi.execute_pauli(wi.lower_pauli(sub, fac, n, rank=r, world=world, phys=phys), view, nl)
        if G:
            src = idx.copy()
            for g, l in zip(G, L):
                diff = ((src >> np.uint64(g)) ^ (src >> np.uint64(l))) & np.uint64(1)
                src ^= (diff << np.uint64(g)) | (diff << np.uint64(l))
            V = V[src]
            exchanges += 1
    return wi.to_logical(V, n, final_phys), exchanges


@pytest.mark.parametrize("n,world,steps", [(11, 2, 2), (12, 4, 2), (13, 8, 2)])
def test_sharded_trotter_replayed_on_cpu(ref, n, world, steps):
    """First-order Trotter steps of the Heisenberg chain on a sharded state: staged around the exchanges with the exact
    Pauli commutation test (about one exchange per step); every rank's fused Pauli-window programs, replayed on the CPU,
    track the oracle's term-by-term evolution."""
    import quant_iron_b200 as gpu
    dt = 0.01
    hg, hr = gpu.heisenberg_1d(n, 1.0, 2.0, 3.0, 0.5, 0.1), ref.heisenberg_1d(n, 1.0, 2.0, 3.0, 0.5, 0.1)
    start = ref.random_state(n, 21)
    want = vec(ref.trotter_evolve_state(hr, start, dt, steps, ref.TrotterOrder.First))
    strings = list(hg.terms) * steps
    got, exchanges = _emulate_pauli(strings, [complex(0.0, -dt)] * len(strings), n, world, start.state_vector)
    assert float(np.max(np.abs(got - want))) <= AMP_TOL
    assert 1 <= exchanges <= 2 * steps + 1


@pytest.mark.parametrize("n,world,seed", [(11, 2, 1), (12, 4, 2), (13, 8, 3), (12, 8, 4)])
def test_sharded_random_pauli_sequence_replayed_on_cpu(ref, n, world, seed):
    from test_window_lowering import _random_strings
    import quant_iron_b200 as gpu
    (sg, sr), factors = _random_strings([gpu, ref], n, 90 + seed, 50)
    r = ref.random_state(n, 33 + seed)
    start = np.array(r.state_vector)
    for p, f in zip(sr, factors):
        r = p.apply_exp_factor(r, f)
    got, exchanges = _emulate_pauli(sg, factors, n, world, start)
    nrm = max(1.0, float(np.sqrt(np.vdot(vec(r), vec(r)).real)))
    assert float(np.max(np.abs(got - vec(r)))) <= AMP_TOL * nrm
    assert exchanges >= 1
