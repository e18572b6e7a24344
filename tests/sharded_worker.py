"""Worker for the multi-rank tests (launched with torchrun, one process per GPU).

    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/sharded_worker.py [--cpu]

--cpu: host-logic only over gloo (no CUDA): handle exchange, un-permutation, planner agreement.
Without it: sharded engine vs the CPU oracle on small states (amplitudes <= 1e-12, expectation 1e-10 relative,
identical seeded bins on every rank), then a larger closed-form QFT check.
"""
import argparse
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cpu", action="store_true")
ap.add_argument("--big", type=int, default=0, help="local qubits for the large closed-form QFT check (0 = skip)")
args = ap.parse_args()

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local_rank = int(os.environ.get("LOCAL_RANK", rank))

import quant_iron_b200 as qi  # noqa: E402
from quant_iron_b200 import sharded, workloads as w  # noqa: E402


def log(*a):
    if rank == 0:
        print(*a, flush=True)


if args.cpu:
    dist.init_process_group("gloo")
    got = sharded.exchange_handles(bytes([rank]) * sharded.HANDLE_BYTES, dist)
    assert [g[0] for g in got] == list(range(world)) and all(len(g) == sharded.HANDLE_BYTES for g in got)
    # un-permutation: a layout that swaps qubits 0 and 3 of a 4-qubit register
    phys = [3, 1, 2, 0]
    physical = np.arange(16, dtype=np.complex128)
    logical = sharded.unpermute(physical, phys)
    for i in range(16):
        p = sum(((i >> q) & 1) << phys[q] for q in range(4))
        assert logical[i] == physical[p]
    # every rank plans the same exchanges (the engine's decisions depend on the gate list only)
    n = 14
    c = w.build_circuit(qi, n, w.random_layered_circuit(n, 6) + w.qft_specs(n))
    pl = sharded.plan(n, world, c)
    plans = [None] * world
    dist.all_gather_object(plans, pl)
    assert all(p == plans[0] for p in plans), plans
    assert sharded.plan(n, 1, c)["exchanges"] == 0
    log(f"PASS cpu host-logic world={world} exchanges={pl['exchanges']} comm_free={pl['comm_free_global_gates']}")
    dist.destroy_process_group()
    sys.exit(0)

torch.cuda.set_device(local_rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
qi.engine.init(local_rank)
from oracle import refapi as ref  # noqa: E402  (checker)

AMP_TOL = 1e-12


def check(name, st, ref_state):
    got = sharded.gather_state_vector(st, dist)
    err = float(np.max(np.abs(got - ref_state.state_vector)))
    assert err <= AMP_TOL, f"{name}: max abs amplitude error {err:.3e}"
    log(f"PASS {name}: max|err|={err:.2e} exchanges={sharded.comm_stats(st)['exchanges']}")


for n in (10, 13):
    # 1. layered circuit touching the global qubits, then QFT
    specs = w.random_layered_circuit(n, 5) + w.qft_specs(n)
    st = sharded.new_zero(n, dist)
    w.build_circuit(qi, n, specs).execute_(st)
    rs = w.build_circuit(ref, n, specs).execute(ref.State.new_zero(n))
    check(f"layered+qft n={n} world={world}", st, rs)
    nrm = st.norm_sqr()
    assert abs(nrm - 1.0) < 1e-12, nrm
    # 2. amplitudes through the layout map, on every rank
    for i in (0, 5, (1 << n) - 1, 1 << (n - 1)):
        assert abs(st.amplitude(i) - complex(rs.state_vector[i])) <= AMP_TOL
    # 3. Pauli exp / expectation with X,Y,Z on global qubits
    hg, hr = qi.heisenberg_1d(n, 1.0, 2.0, 3.0, 0.5, 0.1), ref.heisenberg_1d(n, 1.0, 2.0, 3.0, 0.5, 0.1)
    qi.trotter_evolve_state_(hg, st, 0.01, 2, qi.TrotterOrder.First)
    rs = ref.trotter_evolve_state(hr, rs, 0.01, 2, ref.TrotterOrder.First)
    check(f"trotter n={n}", st, rs)
    eg, er = hg.expectation_value(st), hr.expectation_value(rs)
    assert abs(eg - er) <= 1e-10 * abs(er), (eg, er)
    log(f"PASS expectation n={n}: {eg.real:.12f} vs {er.real:.12f}")
    # 4. controlled gates with controls/targets on global qubits, swaps (relabelled), toffoli
    st.cnot_(n - 1, 0).cnot_(0, n - 1).swap_(1, n - 1).toffoli_(n - 1, n - 2, 2).h_(n - 1).rz_(n - 2, 0.3)
    st.apply_(qi.PhaseShift(0.7), [n - 1], [n - 2]).apply_(qi.RotateX(0.4), [n - 2], [0, n - 1])
    rs = rs.cnot(n - 1, 0).cnot(0, n - 1).swap(1, n - 1).toffoli(n - 1, n - 2, 2).h(n - 1).rz(n - 2, 0.3)
    rs = rs.cp_multi([n - 1], [n - 2], 0.7).crx_multi([n - 2], [0, n - 1], 0.4)
    check(f"global controls/targets n={n}", st, rs)
    # 5. probabilities / seeded sampling / measurement incl. global qubits: identical on all ranks and to the oracle
    qubits = [0, n - 1, 3, n - 2]
    pg, pr = st.probabilities(qubits), rs.probabilities(qubits)
    assert np.max(np.abs(pg - pr)) <= 1e-13
    bins_g, bins_r = st.sample_counts(qubits, 200, 20260003), rs.sample_counts(qubits, 200, 20260003)
    assert np.array_equal(bins_g, bins_r)
    idx, out = st.measure_(qi.MeasurementBasis.Computational, qubits, seed=5)
    mr = rs.measure(ref.MeasurementBasis.Computational, qubits, seed=5)
    assert out == mr.get_outcomes(), (out, mr.get_outcomes())
    check(f"measure n={n}", st, mr.get_new_state())
    del st

# 6. the CTA-tile executor on shards, interpreted (k_tile) and as JIT modules: rank-bit controls are per-rank constants, so
#    every rank assembles its own modules
for jit in (0, 2):
    qi.engine.set_option("tile_min_qubits", 11)
    qi.engine.set_option("jit", jit)
    n = 12 + int(math.log2(world))
    specs = w.random_layered_circuit(n, 8, seed=777) + w.qft_specs(n) + w.random_layered_circuit(n, 3, seed=778)
    st = sharded.new_zero(n, dist)
    qi.engine.stats_reset()
    w.build_circuit(qi, n, specs).execute_(st)
    kernels = {k: v["launches"] for k, v in qi.engine.stats().items()}
    rs = w.build_circuit(ref, n, specs).execute(ref.State.new_zero(n))
    check(f"tile executor on shards jit={jit} n={n} kernels={kernels}", st, rs)
    assert kernels.get("gate_tile_jit" if jit else "gate_tile", 0) > 0, kernels
    del st
qi.engine.set_option("tile_min_qubits", 18)
qi.engine.set_option("jit", 1)

if args.big:
    n = args.big + int(math.log2(world))
    st = sharded.new_plus(n, dist)
    qft = qi.CircuitBuilder(n).add_subroutine(qi.Subroutine.qft(list(range(n)), n)).build()
    qi.engine.synchronize()
    dist.barrier()
    qi.engine.timer_start()
    qft.execute_(st)
    ms = qi.engine.timer_stop()
    a0 = st.amplitude(0)
    nrm = st.norm_sqr()
    assert abs(a0 - 1.0) < 1e-12 and abs(nrm - 1.0) < 1e-10, (a0, nrm)
    cs = sharded.comm_stats(st)
    log(f"PASS qft closed form n={n} world={world}: {ms:.1f} ms, |amp0-1|={abs(a0 - 1):.1e}, exchanges={cs['exchanges']}, "
        f"bytes_sent/rank={cs['bytes_sent']}")
    # Trotter evolution at the same size: fused Pauli-exp passes between the exchanges that bring the chain's
    # far end into the local bits; size-independent checks: norm 1, energy stays at <+|H|+> = -n/2 up to
    # the Trotter error
    del st
    st = sharded.new_plus(n, dist)
    h = qi.heisenberg_1d(n, 1.0, 2.0, 3.0, 0.5, 0.1)
    qi.engine.synchronize()
    dist.barrier()
    qi.engine.timer_start()
    qi.trotter_evolve_state_(h, st, 0.01, 2, qi.TrotterOrder.First)
    ms = qi.engine.timer_stop()
    e = h.expectation_value(st)
    nrm = st.norm_sqr()
    assert abs(nrm - 1.0) < 1e-10 and abs(e.real + 0.5 * n) < 1e-3 * n and abs(e.imag) < 1e-9, (nrm, e)
    cs = sharded.comm_stats(st)
    log(f"PASS trotter energy n={n} world={world}: {ms:.1f} ms for 2 steps, <H>={e.real:.9f}, exchanges={cs['exchanges']}, "
        f"kernels={ {k: v['launches'] for k, v in qi.engine.stats().items() if k.startswith('pauli')} }")
dist.barrier()
log("ALL PASS")
dist.destroy_process_group()
