"""BASELINE.json configs 1, 3, 4 (single GPU) and 5 (torchrun, 2/4/8 GPUs): wall time + closed-form checks.

  python tools/bench_configs.py --configs 1,3,4
  torchrun --nproc-per-node 8 ... tools/bench_configs.py --configs 5 [--local-qubits 33]
Prints one JSON line per config (rank 0)."""
import argparse
import json
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quant_iron_b200 as qi  # noqa: E402
from quant_iron_b200 import workloads as w  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--configs", default="1,3,4")
ap.add_argument("--local-qubits", type=int, default=33)
ap.add_argument("--qft4-qubits", type=int, default=33)
ap.add_argument("--with-oracle", action="store_true", help="config 1/3: also run the CPU oracle and report parity")
a = ap.parse_args()
configs = [int(x) for x in a.configs.split(",")]
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
dist = None
if world > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
qi.engine.init(local_rank)


def emit(d):
    if rank == 0:
        print(json.dumps(d), flush=True)


def timed(fn, reps=1):
    qi.engine.synchronize()
    if dist is not None:
        dist.barrier()
    qi.engine.timer_start()
    for _ in range(reps):
        fn()
    return qi.engine.timer_stop() / reps


if 1 in configs:
    n = 20
    qft = qi.CircuitBuilder(n).add_subroutine(qi.Subroutine.qft(list(range(n)), n)).build()
    st = qi.State.new_plus(n)
    out = qft.execute(st)
    ms = timed(lambda: qft.execute(st), reps=20)
    v = out.state_vector
    rec = {"config": 1, "what": "20-qubit QFT (Subroutine::qft) on new_plus(20) via CircuitBuilder", "gates": len(qft.gates),
           "gpu_ms": ms, "amp0_minus_1": abs(v[0] - 1.0), "max_other_amp": float(abs(v[1:]).max())}
    if a.with_oracle:
        from oracle import refapi as ref
        rq = ref.CircuitBuilder(n).add_subroutine(ref.Subroutine.qft(list(range(n)), n)).build()
        t0 = time.perf_counter()
        ro = rq.execute(ref.State.new_plus(n))
        rec["cpu_oracle_inplace_ms"] = (time.perf_counter() - t0) * 1e3
        rec["cpu_threads"] = ref.num_threads()
        rec["max_abs_amp_err_vs_oracle"] = float(abs(v - ro.state_vector).max())
    emit(rec)

if 3 in configs:
    n = 24
    h = qi.heisenberg_1d(n, 1.0, 2.0, 3.0, 0.5, 0.1)
    st = qi.State.new_plus(n)
    qi.trotter_evolve_state_(h, st, 0.01, 1, qi.TrotterOrder.First)      # warm-up step
    st = qi.State.new_plus(n)
    ms = timed(lambda: qi.trotter_evolve_state_(h, st, 0.01, 50, qi.TrotterOrder.First))
    h.expectation_value(st)                                                # warm-up (scratch allocation, module load)
    t0 = time.perf_counter()
    e = h.expectation_value(st)
    ms_e = (time.perf_counter() - t0) * 1e3
    qi.engine.set_option("fuse", 0)
    t0 = time.perf_counter()
    e_u = h.expectation_value(st)
    ms_e_u = (time.perf_counter() - t0) * 1e3
    qi.engine.set_option("fuse", 1)
    rec = {"config": 3, "what": "heisenberg_1d(24,1,2,3,0.5,0.1), new_plus(24), 50 first-order Trotter steps dt=0.01, expectation",
           "terms": h.num_terms(), "exp_applications": 50 * h.num_terms(), "trotter_ms": ms, "expectation_ms": ms_e, "expectation_ms_per_term_kernels": ms_e_u,
           "expectation_fused_minus_per_term": abs(e - e_u),
           "expectation": [e.real, e.imag], "norm_sqr": st.norm_sqr(),
           "effective_gbs": 50 * h.num_terms() * 32.0 * (1 << n) / (ms * 1e-3) / 1e9}
    # the same evolution on the per-term kernels (one pass per exp), and the pass statistics of the fused path
    qi.engine.set_option("fuse", 0)
    st_u = qi.State.new_plus(n)
    rec["trotter_ms_per_term_kernels"] = timed(lambda: qi.trotter_evolve_state_(h, st_u, 0.01, 50, qi.TrotterOrder.First))
    qi.engine.set_option("fuse", 1)
    rec["fused_vs_per_term_diff_norm"] = (st - st_u).norm_sqr() ** 0.5
    del st_u
    qi.engine.stats_reset()
    qi.engine.set_option("profile", 1)
    st_p = qi.State.new_plus(n)
    qi.trotter_evolve_state_(h, st_p, 0.01, 50, qi.TrotterOrder.First)
    qi.engine.synchronize()
    qi.engine.set_option("profile", 0)
    rec["kernels"] = qi.engine.stats()
    w = rec["kernels"].get("pauli_exp_window")
    if w and w["launches"]:
        rec["window_pass_ms"] = w["total_ms"] / w["launches"]
        rec["window_pass_gbs"] = 32.0 * (1 << n) / (w["total_ms"] / w["launches"] * 1e-3) / 1e9
    del st_p
    if a.with_oracle:
        from oracle import refapi as ref
        hr = ref.heisenberg_1d(n, 1.0, 2.0, 3.0, 0.5, 0.1)
        t0 = time.perf_counter()
        rs = ref.trotter_evolve_state(hr, ref.State.new_plus(n), 0.01, 2, ref.TrotterOrder.First)
        rec["cpu_oracle_ms_per_step"] = (time.perf_counter() - t0) * 1e3 / 2
        st2 = qi.State.new_plus(n)
        qi.trotter_evolve_state_(h, st2, 0.01, 2, qi.TrotterOrder.First)
        er, eg = hr.expectation_value(rs), h.expectation_value(st2)
        rec["expectation_rel_err_vs_oracle_after_2_steps"] = abs(eg - er) / abs(er)
        rec["max_abs_amp_err_vs_oracle_after_2_steps"] = float(abs(st2.state_vector - rs.state_vector).max())
    emit(rec)

if 4 in configs:
    n = a.qft4_qubits
    qft = qi.CircuitBuilder(n).add_subroutine(qi.Subroutine.qft(list(range(n)), n)).build()
    st = qi.State.new_plus(n)
    qi.engine.stats_reset()
    ms = timed(lambda: qft.execute_(st))
    a0 = st.amplitude(0)
    rec = {"config": 4, "what": f"{n}-qubit QFT f64 on one B200 ({16 * (1 << n) / 2**30:.0f} GiB state), new_plus -> |0..0>",
           "gates": len(qft.gates), "gpu_ms": ms, "amp0_minus_1": abs(a0 - 1.0), "norm_sqr": st.norm_sqr(),
           "probe_amps": [abs(st.amplitude(i)) for i in (1, 12345, (1 << n) - 1)],
           "kernels": {k: v["launches"] for k, v in qi.engine.stats().items()}}
    del st
    emit(rec)

if 5 in configs:
    from quant_iron_b200 import sharded
    n = a.local_qubits + int(math.log2(world))
    # (a) pure QFT on |+..+> -> |0..0> (closed form)
    st = sharded.new_plus(n, dist)
    qft = qi.CircuitBuilder(n).add_subroutine(qi.Subroutine.qft(list(range(n)), n)).build()
    qi.engine.stats_reset()
    ms = timed(lambda: qft.execute_(st))
    a0, nrm = st.amplitude(0), st.norm_sqr()
    cs = sharded.comm_stats(st)
    stats = qi.engine.stats()
    rec = {"config": 5, "what": f"{n}-qubit QFT sharded over {world} GPUs ({a.local_qubits} local qubits, "
                                f"{16 * (1 << a.local_qubits) / 2**30:.0f} GiB per GPU)", "gates": len(qft.gates), "qft_ms": ms,
           "amp0_minus_1": abs(a0 - 1.0), "norm_sqr": nrm, "exchanges": cs["exchanges"],
           "bytes_sent_per_rank": cs["bytes_sent"], "kernels": {k: v["launches"] for k, v in stats.items()}}
    # (b) random layered prefix (depth 4, touches the global qubits) + QFT; norm must stay 1
    del st
    st = sharded.new_zero(n, dist)
    pre = w.build_circuit(qi, n, w.random_layered_circuit(n, 4) + w.qft_specs(n))
    qi.engine.set_option("profile", 1)
    qi.engine.stats_reset()
    ms2 = timed(lambda: pre.execute_(st))
    prof = qi.engine.stats()
    qi.engine.set_option("profile", 0)
    cs2 = sharded.comm_stats(st)
    ex_ms = prof.get("exchange", {}).get("total_ms", 0.0)
    rec.update({"prefix_plus_qft_ms": ms2, "prefix_gates": len(pre.gates), "prefix_norm_sqr": st.norm_sqr(),
                "prefix_exchanges": cs2["exchanges"], "exchange_ms_total": ex_ms,
                "nvlink_gbs_per_gpu_per_direction": cs2["bytes_sent"] / max(1e-9, ex_ms * 1e-3) / 1e9,
                "per_kernel_ms": {k: round(v["total_ms"], 2) for k, v in prof.items()}})
    emit(rec)
    del st
if dist is not None:
    dist.barrier()
    dist.destroy_process_group()
