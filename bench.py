#!/usr/bin/env python
"""bench.py -- headline benchmark of the state-vector hot path (contract in the task statement).

  python bench.py --gpus N --steps K --warmup W            # this build (B200)
  python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path

Metric (BASELINE.json): gates/s on the synthetic random layered circuit (H/RX/RZ/CNOT, depth 40,
configs[1]'s generator) at the metric's 30 qubits per GPU, f64, plus the achieved HBM GB/s of the
dominant kernel against the measured roofline, plus single-gate passes at 30 qubits.
A "step" is one execution of the whole circuit on a state that is already resident in HBM; `e2e`
is the same circuit through the public API with HOST buffers (H2D of the initial state and D2H of the
final state inside the timed region).  At N = 1 `e2e` goes through Circuit.execute_host_ (qi_execute_host:
chunked copies overlapped with the circuit) once that entry has reproduced the plain upload / execute /
download sequence on this box; `e2e.mode` says which one the value is, `e2e.serial_value` is the plain
sequence either way.  N > 1: the state is sharded (top log2 N qubits global),
30 local qubits per GPU (weak scaling).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "gates_per_sec"
UNIT = "gates/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------
def load_workloads():
    """quant_iron_b200/workloads.py by PATH: the generator is pure Python, and the reference arm must not map the
    product's library into its process (importing the package would)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_qi_workloads", os.path.join(ROOT, "quant_iron_b200", "workloads.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod
    spec.loader.exec_module(mod)
    return mod


def host_info():
    cpu_model = None
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    cpu_model = line.split(":", 1)[1].strip()
                    break
    except OSError:
        pass
    try:
        avail = int(open("/proc/meminfo").read().split("MemAvailable:")[1].split()[0]) * 1024
    except Exception:
        avail = 32 << 30
    return cpu_model, avail


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm runs on rank 0 alone and uses every host core, so the
    variable is set explicitly BEFORE the oracle's OpenMP runtime is loaded."""
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)


def cpu_baseline(num_qubits: int, depth: int, budget_s: float = 20.0):
    """The reference's CPU path, restated (oracle/qi_oracle.c), timed on this box's host cores on a bounded sample of the
    SAME workload: whole layers (every single-qubit gate of the layer plus its CNOT row; at least one full layer) of the
    same generator at the largest qubit count whose layer fits the time budget and host memory.
    `faithful` = the reference's rayon-branch pass structure (clone + parallel (index, value) updates with a heap
    allocation per pair + serial scatter, operator.rs:339-360); `inplace` = the same arithmetic in place (a stronger CPU
    baseline).  A gate's cost is proportional to the state size, so the rate at the bench size is the measured rate times
    2^-(bench qubits - sample qubits); both numbers and the factor are reported."""
    import numpy as np
    from oracle import refapi as ref
    w = load_workloads()
    kinds = {"h": ref.G_H, "rx": ref.G_RX, "rz": ref.G_RZ, "cnot": ref.G_CNOT}
    cores = ref.num_threads()
    cpu_model, avail = host_info()
    # calibrate both variants at 22 qubits (one H)
    n_cal = min(22, num_qubits)
    v = np.zeros(1 << n_cal, dtype=np.complex128)
    v[0] = 1.0
    dst = np.empty_like(v)
    ref.gate_faithful(v, dst, n_cal, ref.G_H, [n_cal // 2], [], [])
    t0 = time.perf_counter()
    ref.gate_faithful(v, dst, n_cal, ref.G_H, [n_cal // 2], [], [])
    per_amp_f = (time.perf_counter() - t0) / float(1 << n_cal)
    t0 = time.perf_counter()
    ref.gate_inplace(v, n_cal, ref.G_H, [n_cal // 2], [], [])
    per_amp_i = (time.perf_counter() - t0) / float(1 << n_cal)
    del v, dst

    def pick(per_amp, share, bytes_per_amp):
        n = min(num_qubits, 30)
        while n > n_cal and (per_amp * (1 << n) * 1.5 * n > budget_s * share or bytes_per_amp * (1 << n) > avail * 0.6):
            n -= 1
        return n

    def run(kind, n, share):
        specs = w.random_layered_circuit(n, depth)
        per_layer = n + (n - 1 + 1) // 2                  # upper bound; layers are delimited by counting single-qubit gates
        a = np.zeros(1 << n, dtype=np.complex128)
        a[0] = 1.0
        b = np.empty_like(a) if kind == "faithful" else None
        done = layers = in_layer_1q = 0
        t_start = time.perf_counter()
        t_layer_end = t_start
        i = 0
        while i < len(specs):
            # one whole layer: n single-qubit gates, then the CNOT row
            j = i
            one_q = 0
            while j < len(specs) and (one_q < n or specs[j][0] == "cnot"):
                one_q += specs[j][0] != "cnot"
                j += 1
            for name, targets, controls, params in specs[i:j]:
                if kind == "faithful":
                    ref.gate_faithful(a, b, n, kinds[name], targets, controls, params)
                    a, b = b, a
                else:
                    ref.gate_inplace(a, n, kinds[name], targets, controls, params)
            done += j - i
            layers += 1
            i = j
            t_layer_end = time.perf_counter()
            if t_layer_end - t_start > budget_s * share:
                break
        del per_layer, in_layer_1q
        dt = t_layer_end - t_start
        return {"qubits": n, "layers": layers, "gates": done, "seconds": dt, "gates_per_sec_at_sample": done / dt}

    f = run("faithful", pick(per_amp_f, 0.6, 100), 0.6)
    g = run("inplace", pick(per_amp_i, 0.3, 20), 0.3)
    sf, sg = float(1 << (num_qubits - f["qubits"])), float(1 << (num_qubits - g["qubits"]))
    return {
        "cpu_model": cpu_model, "nproc": os.cpu_count(),
        "value": f["gates_per_sec_at_sample"] / sf, "unit": UNIT, "cores": cores, "kind": "port",
        "sample": f"{f['layers']} whole layer(s) = {f['gates']} gates (every single-qubit gate + the CNOT row) of the same generator at "
                  f"{f['qubits']} qubits, reference pass structure (clone + per-pair updates + serial scatter, operator.rs:339-360), "
                  f"{f['seconds']:.1f} s measured; rate scaled by 2^-{num_qubits - f['qubits']} to the bench size",
        "faithful": f, "inplace": g, "scale_factor_faithful": sf, "scale_factor_inplace": sg,
        "scaled_to_bench_qubits": f["gates_per_sec_at_sample"] / sf,
        "inplace_scaled_to_bench_qubits": g["gates_per_sec_at_sample"] / sg,
    }


def run_reference(args, rank, world):
    """`--impl reference`: the reference's own CPU implementation of the path (its Rust crate cannot be built here; the C
    restatement under oracle/ is what is timed), all host threads, rank 0 only.  A step = a bounded sample of the workload
    (whole layers at the largest size that fits); `ms_per_step` is the MEASURED time of that sample, `value` the measured
    rate scaled to the bench size by the stated factor."""
    if rank != 0:
        return
    use_all_host_threads()
    n = args.qubits + (world.bit_length() - 1)       # same total size as the GPU arm at this N
    vals, times, base = [], [], None
    budget = min(args.cpu_budget, 150.0 / (args.warmup + args.steps))   # whole run ends within a few minutes
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        base = cpu_baseline(n, args.depth, budget_s=budget)
        if i >= args.warmup:
            vals.append(base["scaled_to_bench_qubits"] * world)     # unit of work: gate x 2^n_per_gpu shard (see config)
            times.append(time.perf_counter() - t0)
    value = sum(vals) / len(vals)
    f = base["faithful"]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"random layered circuit H/RX/RZ/CNOT depth {args.depth}, {n} qubits; each step times a bounded sample: "
                               f"{f['layers']} whole layer(s) at {f['qubits']} qubits (reference pass structure) and "
                               f"{base['inplace']['layers']} layer(s) at {base['inplace']['qubits']} qubits (in-place port); "
                               f"value = measured rate x 2^-{n - f['qubits']} (x N: one unit of work = one gate on one 2^{args.qubits} shard)",
                   "sample_qubits": f["qubits"], "scale_factor": base["scale_factor_faithful"],
                   "ms_per_step_is": "measured wall time of one step's samples (both variants), not an extrapolation",
                   "inplace_port_value": base["inplace_scaled_to_bench_qubits"] * world},
        "cpu_baseline": dict(base, value=value),
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------
def single_gate_table(qi, n, peak_gbs, passes=20):
    """30-qubit single-gate microbenchmark: gate classes {H, RX, RZ, P, X, CNOT, CP, U2} x targets
    {0,1,2,3,4,5,8,12,16,20,24,28,29}; `passes` back-to-back passes per cell after one warm-up pass, timed with
    CUDA events on the engine stream; achieved GB/s = algorithmic bytes (2*16*2^n*f) / mean pass time."""
    import math
    st = qi.State.new_random(n)
    full = 2.0 * 16.0 * float(1 << n)
    ctrl = n - 2
    c, s_ = math.cos(0.35), math.sin(0.35)
    u2 = qi.Unitary2.new([[complex(c, 0.0), complex(0.0, -s_) * complex(math.cos(0.2), math.sin(0.2))],
                          [complex(0.0, -s_) * complex(math.cos(0.2), -math.sin(0.2)), complex(c, 0.0)]])
    gates = {
        "h": (lambda t: st.h_(t), 1.0), "rx": (lambda t: st.rx_(t, 0.3), 1.0), "rz": (lambda t: st.rz_(t, 0.3), 1.0),
        "u2": (lambda t: st.apply_(u2, [t], []), 1.0),
        "x": (lambda t: st.x_(t), 1.0), "p": (lambda t: st.p_(t, 0.3), 0.5),
        "cnot": (lambda t: st.cnot_(ctrl if t != ctrl else ctrl - 1, t), 0.5),
        "cp": (lambda t: st.cp_multi_([t], [ctrl if t != ctrl else ctrl - 1], 0.3), 0.25),
    }
    targets = sorted(set(t for t in (0, 1, 2, 3, 4, 5, 8, 12, 16, 20, 24, 28, 29) if t < n) | {n - 1})
    table = {}
    for name, (fn, frac) in gates.items():
        row = {}
        for t in targets:
            fn(t)
            qi.engine.synchronize()
            qi.engine.timer_start()
            for _ in range(passes):
                fn(t)
            ms = qi.engine.timer_stop() / passes
            row[str(t)] = round(full * frac / (ms * 1e-3) / 1e9, 1)
        table[name] = row
    best_h = max(table["h"].values())
    worst_h = min(table["h"].values())
    full_f = [v for k in ("h", "rx", "rz", "u2", "x") for v in table[k].values()]
    return {"qubits": n, "passes": passes, "gbs": table, "h_best_frac_of_measured_peak": best_h / peak_gbs,
            "h_worst_frac_of_measured_peak": worst_h / peak_gbs, "h_best_frac_of_8TBs_nominal": best_h / 8000.0,
            "f1_gates_min_frac_of_8TBs_nominal": min(full_f) / 8000.0, "f1_gates_min_frac_of_measured_peak": min(full_f) / peak_gbs}


def _kernel_rooflines(prof, peak_gbs):
    """achieved algorithmic GB/s per kernel family of a profiled run (qi_set_option("profile", 1) brackets every launch with
    CUDA events on the engine stream; algorithmic bytes per launch are the library's own accounting)."""
    out = {}
    for k, v in prof.items():
        if v["launches"] and v["total_ms"] > 0 and v["algorithmic_bytes"] > 0:
            gbs = v["algorithmic_bytes"] / (v["total_ms"] * 1e-3) / 1e9
            out[k] = {"launches": v["launches"], "avg_launch_ms": round(v["total_ms"] / v["launches"], 4), "achieved_gbs": round(gbs, 1),
                      "frac_of_measured_hbm": round(gbs / peak_gbs, 3)}
    return out


def probe_indices(n, count=64, seed=20260007):
    """`count` amplitude indices below 2^n from the splitmix64 stream (the same on every rank)."""
    from quant_iron_b200 import workloads as w
    rng = w.SplitMix64(seed)
    return [rng.next_u64() & ((1 << n) - 1) for _ in range(count)]


def extras_single_gpu(qi, w, peak_gbs):
    """BASELINE configs 1, 3, 4 and the measurement path on one B200, each with a correctness field next to its time."""
    import numpy as np
    out = {}
    # ---- config 1: 20-qubit QFT (Subroutine::qft, subroutine.rs:90-112) on new_plus(20) via CircuitBuilder; full parity vs the oracle
    try:
        n = 20
        qft = qi.CircuitBuilder(n).add_subroutine(qi.Subroutine.qft(list(range(n)), n)).build()
        st = qi.State.new_plus(n)
        res = qft.execute(st)
        qi.engine.synchronize()
        qi.engine.timer_start()
        for _ in range(20):
            qft.execute(st)
        ms_clone = qi.engine.timer_stop() / 20
        work = qi.State.new_plus(n)
        qi.engine.timer_start()
        for _ in range(20):
            qft.execute_(work)
        ms_inplace = qi.engine.timer_stop() / 20
        v = res.state_vector
        rec = {"gates": len(qft.gates), "gpu_ms_execute": ms_clone, "gpu_ms_execute_in_place": ms_inplace,
               "amp0_minus_1": abs(v[0] - 1.0), "max_other_amp": float(np.abs(v[1:]).max())}
        from oracle import refapi as ref                       # checker + CPU time of the same call (cpu_baseline leg)
        rq = ref.CircuitBuilder(n).add_subroutine(ref.Subroutine.qft(list(range(n)), n)).build()
        t0 = time.perf_counter()
        ro = rq.execute(ref.State.new_plus(n))
        rec["cpu_oracle_inplace_ms"] = (time.perf_counter() - t0) * 1e3
        rec["cpu_threads"] = ref.num_threads()
        rec["max_abs_amp_err_vs_oracle"] = float(np.abs(v - ro.state_vector).max())
        out["config1_qft20"] = rec
        del st, res, work
    except Exception as ex:  # noqa: BLE001
        out["config1_qft20"] = {"error": repr(ex)[:300]}
    # ---- config 3: heisenberg_1d(24), 50 first-order Trotter steps (time_evolution.rs:140-167) + <H> (pauli_string.rs:485-507)
    try:
        n = 24
        h = qi.heisenberg_1d(n, 1.0, 2.0, 3.0, 0.5, 0.1)
        st = qi.State.new_plus(n)
        qi.trotter_evolve_state_(h, st, 0.01, 1, qi.TrotterOrder.First)
        st = qi.State.new_plus(n)
        qi.engine.synchronize()
        qi.engine.timer_start()
        qi.trotter_evolve_state_(h, st, 0.01, 50, qi.TrotterOrder.First)
        ms = qi.engine.timer_stop()
        h.expectation_value(st)
        qi.engine.synchronize()
        t0 = time.perf_counter()
        e = h.expectation_value(st)
        ms_e = (time.perf_counter() - t0) * 1e3
        rec = {"terms": h.num_terms(), "exp_applications": 50 * h.num_terms(), "trotter_ms": ms, "expectation_ms": ms_e,
               "expectation": [e.real, e.imag], "norm_sqr": st.norm_sqr()}
        qi.engine.stats_reset()
        qi.engine.set_option("profile", 1)
        sp = qi.State.new_plus(n)
        qi.trotter_evolve_state_(h, sp, 0.01, 50, qi.TrotterOrder.First)
        h.expectation_value(sp)
        qi.engine.synchronize()
        qi.engine.set_option("profile", 0)
        rec["kernels"] = _kernel_rooflines(qi.engine.stats(), peak_gbs)
        del sp, st
        # parity vs the oracle at 16 sites (the 24-site oracle takes minutes; tests/test_gpu_parity.py holds the 24-site case)
        from oracle import refapi as ref
        m = 16
        hg, hr = qi.heisenberg_1d(m, 1.0, 2.0, 3.0, 0.5, 0.1), ref.heisenberg_1d(m, 1.0, 2.0, 3.0, 0.5, 0.1)
        sg = qi.State.new_plus(m)
        qi.trotter_evolve_state_(hg, sg, 0.01, 5, qi.TrotterOrder.First)
        sr = ref.trotter_evolve_state(hr, ref.State.new_plus(m), 0.01, 5, ref.TrotterOrder.First)
        eg, er = hg.expectation_value(sg), hr.expectation_value(sr)
        rec["parity_16_sites_5_steps"] = {"expectation_rel_err": abs(eg - er) / abs(er),
                                          "max_abs_amp_err": float(np.abs(sg.state_vector - sr.state_vector).max())}
        out["config3_heisenberg24_trotter50"] = rec
    except Exception as ex:  # noqa: BLE001
        out["config3_heisenberg24_trotter50"] = {"error": repr(ex)[:300]}
    # ---- measurement (state.rs:525-784) at 30 qubits: marginal probabilities, seeded sampling, collapse
    try:
        n = 30
        st = qi.State.new_random(n)
        qs = [0, 7, 13, 22, 29]
        st.probabilities(qs)
        qi.engine.stats_reset()
        qi.engine.set_option("profile", 1)
        p = st.probabilities(qs)
        bins = st.sample_counts(qs, 4096, 20260003)
        st.measure_(qi.MeasurementBasis.Computational, qs, seed=20260003)
        qi.engine.synchronize()
        qi.engine.set_option("profile", 0)
        out["measurement_30q"] = {"qubits": qs, "prob_sum_minus_1": abs(float(p.sum()) - 1.0), "shots": int(bins.sum()),
                                  "norm_after_collapse": st.norm_sqr(), "kernels": _kernel_rooflines(qi.engine.stats(), peak_gbs)}
        del st
    except Exception as ex:  # noqa: BLE001
        out["measurement_30q"] = {"error": repr(ex)[:300]}
    # ---- config 4: 33-qubit QFT on one B200 (128 GiB state, in place only): closed form QFT|+..+> = |0..0>
    try:
        n = 33
        info = qi.engine.device_info()
        if info["free_mem"] < 16 * (1 << n) + (2 << 30):
            raise RuntimeError(f"needs 128 GiB of free HBM, {info['free_mem'] >> 30} GiB free")
        qft = qi.CircuitBuilder(n).add_subroutine(qi.Subroutine.qft(list(range(n)), n)).build()
        st = qi.State.new_plus(n)

        def run_once():
            qi.engine.stats_reset()
            qi.engine.synchronize()
            qi.engine.timer_start()
            qft.execute_(st)
            return qi.engine.timer_stop(), {k: v["launches"] for k, v in qi.engine.stats().items()}, \
                sum(v["algorithmic_bytes"] for v in qi.engine.stats().values())
        # 1st execution: nothing is assembled yet, the interpreting tile kernel runs every pass.  The QFT's trailing swaps are
        # a relabelling, so the layout -- and with it the pass structures -- alternate with period 2: after two executions and
        # a drain every module exists; executions 3 (|+> -> |0>) and 4 run on them.
        ms1, k1, _ = run_once()
        a0_first = st.amplitude(0)
        run_once()
        qi.engine.jit_drain()
        ms3, k3, bytes3 = run_once()
        a0 = st.amplitude(0)
        rec = {"gates": len(qft.gates), "gpu_ms": ms3, "gpu_ms_first_execution": ms1, "amp0_minus_1": abs(a0 - 1.0),
               "amp0_minus_1_first_execution": abs(a0_first - 1.0), "norm_sqr": st.norm_sqr(),
               "max_probe_amp": max(abs(st.amplitude(i)) for i in probe_indices(n, 16) if i),
               "kernels": k3, "kernels_first_execution": k1,
               "effective_gbs_one_pass_per_launch": bytes3 / (ms3 * 1e-3) / 1e9}
        ms4, _, _ = run_once()
        rec["gpu_ms_4th_execution"] = ms4
        out["config4_qft33"] = rec
        del st
    except Exception as ex:  # noqa: BLE001
        out["config4_qft33"] = {"error": repr(ex)[:300]}
    return out


def extras_multi_gpu(qi, w, dist, world, n_local_cfg5):
    """BASELINE config 5 on N GPUs (QFT on 33 local qubits per GPU: 34/35/36 qubits on 2/4/8) and a sharded-vs-single-GPU
    parity field: the same circuit on the N-GPU sharded state and on one GPU of this rank, 64 seeded probe amplitudes."""
    import math
    from quant_iron_b200 import sharded
    out = {}
    p = int(math.log2(world))

    def timed(fn):
        qi.engine.synchronize()
        dist.barrier()
        qi.engine.timer_start()
        fn()
        return qi.engine.timer_stop()

    try:
        n = n_local_cfg5 + p
        st = sharded.new_plus(n, dist)
        qft = qi.CircuitBuilder(n).add_subroutine(qi.Subroutine.qft(list(range(n)), n)).build()
        # executions 1 and 2 run (partly) on the interpreting tile kernel while every rank assembles its modules; the QFT's
        # trailing swaps relabel the qubits, so the layout alternates with period 2 and execution 3 (|+> -> |0> again) finds
        # every module.  wall_ms is execution 3; the first execution is reported next to it.
        ms_first = timed(lambda: qft.execute_(st))
        a0_first = st.amplitude(0)
        qft.execute_(st)
        qi.engine.jit_drain()
        cs0 = sharded.comm_stats(st)
        qi.engine.stats_reset()
        qi.engine.set_option("profile", 1)
        ms = timed(lambda: qft.execute_(st))
        prof = qi.engine.stats()
        qi.engine.set_option("profile", 0)
        a0, nrm = st.amplitude(0), st.norm_sqr()
        probes = max(abs(st.amplitude(i)) for i in probe_indices(n, 8) if i)
        cs = sharded.comm_stats(st)
        ex_ms = prof.get("exchange", {}).get("total_ms", 0.0)
        sent = cs["bytes_sent"] - cs0["bytes_sent"]
        out[f"config5_qft{n}"] = {
            "qubits": n, "local_qubits": n_local_cfg5, "gib_per_gpu": 16 * (1 << n_local_cfg5) / 2**30, "gates": len(qft.gates),
            "wall_ms": ms, "wall_ms_first_execution": ms_first, "exchanges": cs["exchanges"] - cs0["exchanges"], "bytes_sent_per_rank": sent,
            "exchange_ms": ex_ms, "nvlink_gbs_per_gpu_per_direction": sent / max(1e-9, ex_ms * 1e-3) / 1e9 if ex_ms else None,
            "amp0_minus_1": abs(a0 - 1.0), "amp0_minus_1_first_execution": abs(a0_first - 1.0), "norm_sqr": nrm, "max_probe_amp": probes,
            "per_kernel_ms": {k: round(v["total_ms"], 2) for k, v in prof.items()}, "jit": qi.engine.jit_stats()}
        del st
    except Exception as ex:  # noqa: BLE001
        out["config5_qft"] = {"error": repr(ex)[:300]}
    try:
        n = 30                                                   # <= 33: one GPU holds the whole state next to a shard
        specs = w.random_layered_circuit(n, 8) + w.qft_specs(n)
        circ = w.build_circuit(qi, n, specs)
        st = sharded.new_zero(n, dist)
        circ.execute_(st)
        one = qi.State.new_zero(n)
        circ.execute_(one)
        idx = probe_indices(n)
        diff = max(abs(st.amplitude(i) - one.amplitude(i)) for i in idx)
        out["sharded_vs_single_gpu"] = {"qubits": n, "gates": len(specs), "probes": len(idx), "max_abs_amp_diff": diff,
                                        "norm_sharded": st.norm_sqr(), "norm_single": one.norm_sqr(),
                                        "exchanges": sharded.comm_stats(st)["exchanges"]}
        del st, one
    except Exception as ex:  # noqa: BLE001
        out["sharded_vs_single_gpu"] = {"error": repr(ex)[:300]}
    return out


def run_ours(args, rank, world, local_rank):
    import numpy as np
    import torch
    import quant_iron_b200 as qi
    from quant_iron_b200 import workloads as w

    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    qi.engine.init(local_rank)
    options = {}
    for kv in args.opt:
        name, _, val = kv.partition("=")
        qi.engine.set_option(name, int(val))
        options[name] = int(val)
    peak_gbs, peak_src = load_peaks()
    n_local = args.qubits
    n = n_local + (world.bit_length() - 1)
    specs = w.random_layered_circuit(n, args.depth)
    n_gates = len(specs)

    def barrier():
        qi.engine.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    if world > 1:
        from quant_iron_b200 import sharded
        state = sharded.new_zero(n, dist)
    else:
        state = qi.State.new_zero(n)
    circuit = w.build_circuit(qi, n, specs)

    t_jit0 = time.time()
    jit_wait_s = 0.0
    for i in range(args.warmup):
        circuit.execute_(state)
        if i == 0:
            qi.engine.jit_drain()          # the tile modules of this circuit are assembled during warm-up (background workers, csrc/tile_jit.cuh)
            jit_wait_s = time.time() - t_jit0
    barrier()
    qi.engine.stats_reset()
    sampler = ClockSampler(local_rank)
    sampler.start()
    qi.engine.timer_start()
    for _ in range(args.steps):
        circuit.execute_(state)
    ms_total = qi.engine.timer_stop()
    barrier()
    clocks = sampler.stop()
    stats = qi.engine.stats()
    jit_timed = qi.engine.jit_stats()
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    # unit of work: one gate applied to one 2^n_local-amplitude shard; a gate on the N-GPU state is N units
    # (weak scaling: the per-GPU work of a gate is the same at every N)
    raw_gates_per_sec = n_gates / (ms_per_step * 1e-3)
    value = raw_gates_per_sec * world
    launches = sum(v["launches"] for v in stats.values())

    # roofline of the dominant kernel: per-launch event timing in a separate (profiled) pass
    qi.engine.set_option("profile", 1)
    qi.engine.stats_reset()
    circuit.execute_(state)
    qi.engine.synchronize()
    prof = qi.engine.stats()
    qi.engine.set_option("profile", 0)
    dom = max(prof.items(), key=lambda kv: kv[1]["total_ms"])
    dname, d = dom
    avg_ms = d["total_ms"] / d["launches"]
    bytes_per_launch = d["algorithmic_bytes"] / d["launches"]
    achieved = bytes_per_launch / (avg_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")      # dram bytes per launch from the committed ncu capture
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(dname)
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": dname, "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                "frac": achieved / peak_gbs, "traffic": traffic, "peak_source": peak_src,
                "launches_per_step": d["launches"], "avg_launch_ms": avg_ms,
                "algorithmic_bytes_per_launch": bytes_per_launch,
                "share_of_step": d["total_ms"] / max(1e-9, sum(v["total_ms"] for v in prof.values())),
                "per_kernel_ms": {k: round(v["total_ms"], 3) for k, v in prof.items()}}
    # the tile modules are bound by the FP64 pipe, not by HBM (one pass carries ~100 gates): instruction roofline next to the
    # HBM one.  Numerator: FP64 warp instructions the modules launched in the timed steps (static count per pass, controlled
    # ops weighted by the fraction of threads their controls select; qi_jit_stats).  Denominator: the DMUL+DFMA issue rate
    # measured on this GPU type by tools/micro/fp64_peak.cu (profiles/r02_fp64_dmma_peak.txt: 1.936 warp-instr/clk/SM at the
    # nominal 1965 MHz over 148 SMs).
    if jit_timed["fp64_warp_instr"] > 0 and dname == "gate_tile_jit":
        peak_rate = 1.936 * 148 * 1.965e9
        rate = jit_timed["fp64_warp_instr"] / (ms_total * 1e-3)
        roofline["fp64_pipe"] = {"achieved_warp_instr_per_s": rate, "peak_warp_instr_per_s": peak_rate, "frac": rate / peak_rate,
                                 "fp64_warp_instr_per_step": jit_timed["fp64_warp_instr"] / args.steps,
                                 "peak_source": "measured DMUL+DFMA issue rate, profiles/r02_fp64_dmma_peak.txt",
                                 "note": "the dominant kernel is FP64-issue bound: `frac` above (HBM) is low BECAUSE ~100 gates share one HBM pass"}
    jit_info = dict(qi.engine.jit_stats(), wait_in_warmup_s=round(jit_wait_s, 2))
    jit_info.pop("fp64_warp_instr", None)
    norm = state.norm_sqr()
    comm = None
    if world > 1:
        from quant_iron_b200 import sharded
        ones = torch.ones(1, dtype=torch.float64, device="cuda")
        dist.all_reduce(ones)                                  # NCCL sees every rank: the sum of ones is the rank count
        cs = sharded.comm_stats(state)
        per_step_ex = cs["exchanges"] / (args.warmup + args.steps + 1)
        ex_ms = prof.get("exchange", {}).get("total_ms", 0.0)
        comm = {"nranks": dist.get_world_size(), "nccl_allreduce_of_ones": float(ones.item()), "exchanges_per_step": per_step_ex, "bytes_sent_per_rank_per_step": cs["bytes_sent"] / (args.warmup + args.steps + 1),
                "exchange_ms_per_step": ex_ms,
                "nvlink_gbs_per_gpu_per_direction": (cs["bytes_sent"] / (args.warmup + args.steps + 1)) / max(1e-9, ex_ms * 1e-3) / 1e9}

    # ---- e2e through the public API with HOST buffers: H2D of the initial state from pinned memory,
    # the circuit, D2H of the final state, all inside the timed region; every rank moves its own shard ----
    e2e = None
    try:
        if args.skip_e2e:
            raise RuntimeError("skipped (--skip-e2e)")
        # one pinned buffer per rank, used for both directions (the final state of one step is the initial state
        # of the next: any normalised state times the same); guarded so that N ranks never pin more than the
        # host has
        need = world * 16 * (1 << n_local)
        avail = None
        try:
            with open("/proc/meminfo") as f:
                for line in f:
                    if line.startswith("MemAvailable:"):
                        avail = int(line.split()[1]) * 1024
        except OSError:
            pass
        if avail is not None and need * 1.25 > avail:
            raise RuntimeError(f"host memory: e2e needs {need >> 30} GiB pinned, {avail >> 30} GiB available")
        host_in = torch.zeros(1 << n_local, dtype=torch.complex128).pin_memory()
        host_out = host_in
        if rank == 0:
            host_in[0] = 1.0
        hin, hout = host_in.numpy(), host_out.numpy()
        e2e_steps = max(1, min(args.steps, 3))

        def e2e_step(mode):
            if mode == "pipelined":
                # qi_execute_host: chunked upload / download overlapped with the circuit (csrc/host_pipeline.cu)
                circuit.execute_host_(state, hin, hout)
            else:
                state.upload_(hin)                         # H2D of the initial state (pinned)
                circuit.execute_(state)
                if world > 1:
                    sharded.shard_to_host(state, hout)     # D2H of this rank's shard (physical order)
                else:
                    state.to_host(hout)                    # D2H of the final state

        def e2e_time(mode):
            times = []
            for i in range(1 + e2e_steps):
                barrier()
                t0 = time.perf_counter()
                e2e_step(mode)
                barrier()
                dt = time.perf_counter() - t0
                if i > 0:
                    times.append(dt)
            tt = torch.tensor([sum(times) / len(times)], dtype=torch.float64, device="cuda")
            if dist is not None:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return float(tt.item())

        # the pipelined entry is used for the headline only if it reproduces the plain sequence on this box:
        # both run once from |0..0>, a strided sample of the 2^n outputs (every chunk is hit) must agree to 1e-12
        pipe = {"mode": "serial"}
        if world == 1:
            try:
                def from_zero(mode):
                    host_in.zero_()
                    host_in[0] = 1.0
                    e2e_step(mode)
                    qi.engine.synchronize()
                    return hout[::1021].copy()
                a, b = from_zero("pipelined"), from_zero("serial")
                diff = float(np.max(np.abs(a - b)))
                pipe["pipelined_max_abs_diff_vs_serial"] = diff
                if diff <= 1e-12 and float(np.max(np.abs(b))) > 0.0:
                    pipe["mode"] = "pipelined"
            except Exception as ex:  # noqa: BLE001
                pipe["pipelined_error"] = repr(ex)[:200]
        sec_serial = e2e_time("serial")
        pipe["serial_value"] = n_gates * world / sec_serial
        sec = sec_serial
        if pipe["mode"] == "pipelined":
            try:
                sec = e2e_time("pipelined")
                pipe["pipelined_value"] = n_gates * world / sec
                if sec > sec_serial:             # never report the slower of two verified, equivalent public calls
                    pipe["mode"], sec = "serial", sec_serial
            except Exception as ex:  # noqa: BLE001
                pipe["mode"], pipe["pipelined_error"], sec = "serial", repr(ex)[:200], sec_serial
        e2e_val = n_gates * world / sec
        e2e = {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": world * (16 * (1 << n_local)) + 88 * n_gates,
               "d2h_bytes_per_step": world * 16 * (1 << n_local), "qubits": n, "steps": e2e_steps,
               "seconds_per_step": sec,
               "checksum_norm_first_64k": float(np.vdot(hout[:1 << 16], hout[:1 << 16]).real)}
        e2e.update(pipe)
        del host_in, host_out, hin, hout
    except Exception as ex:  # noqa: BLE001
        e2e = {"value": None, "unit": UNIT, "error": repr(ex)[:200], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    del state

    extras = {}
    if world > 1 and not args.skip_extras:
        extras.update(extras_multi_gpu(qi, w, dist, world, args.config5_local_qubits))      # collective: every rank takes part
    if rank != 0:
        if dist is not None:
            dist.barrier()
        return

    if world == 1 and not args.skip_extras:
        extras.update(extras_single_gpu(qi, w, peak_gbs))
    if world == 1 and not args.skip_extras:
        try:
            extras["single_gate"] = single_gate_table(qi, n_local, peak_gbs)
        except Exception as ex:  # noqa: BLE001
            extras["single_gate"] = {"error": repr(ex)[:200]}
        try:
            n28 = 28
            specs28 = w.random_layered_circuit(n28, 40)
            c28 = w.build_circuit(qi, n28, specs28)
            s28 = qi.State.new_zero(n28)
            qi.engine.synchronize()
            qi.engine.timer_start()
            c28.execute_(s28)                                  # first execution: nothing assembled yet (interpreting tile kernel)
            ms28_first = qi.engine.timer_stop()
            qi.engine.jit_drain()
            qi.engine.stats_reset()
            qi.engine.timer_start()
            for _ in range(3):
                c28.execute_(s28)
            ms28 = qi.engine.timer_stop() / 3
            extras["config_28q_layered_depth40"] = {"gates": len(specs28), "ms_per_circuit": ms28, "ms_first_execution": ms28_first,
                                                    "gates_per_sec": len(specs28) / (ms28 * 1e-3), "norm_sqr": s28.norm_sqr(),
                                                    "kernels": {k: v["launches"] for k, v in qi.engine.stats().items()}}
            del s28
        except Exception as ex:  # noqa: BLE001
            extras["config_28q_layered_depth40"] = {"error": repr(ex)[:200]}

    cpu = None
    if world == 1 and not args.skip_cpu:
        try:
            use_all_host_threads()
            cpu = cpu_baseline(n, args.depth, budget_s=args.cpu_budget)
        except Exception as ex:  # noqa: BLE001
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {ex!r}"[:200]}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"random layered circuit H/RX/RZ/CNOT depth {args.depth} (BASELINE configs[1] generator, "
                               f"seed 20260001) at {n_local} qubits per GPU ({n} qubits total), state resident in HBM",
                   "qubits": n, "gates_per_step": n_gates, "state_bytes_per_gpu": 16 * (1 << n_local),
                   "unit_of_work": "one gate applied to one 2^30-amplitude shard; a gate on the N-GPU state counts N "
                                   "(value = gates/s x N; raw_gates_per_sec is the literal circuit-gate rate)",
                   "raw_gates_per_sec": raw_gates_per_sec, "options": options,
                   "l2": "inputs (16 GiB state per GPU) far larger than the 126 MB L2; no flush needed",
                   "unfused_algorithmic_bytes_per_step": w.algorithmic_bytes(n, specs),
                   "effective_gbs_per_gpu_vs_unfused_bytes": w.algorithmic_bytes(n, specs) / world / (ms_per_step * 1e-3) / 1e9},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches,
        "kernels": {k: v["launches"] for k, v in stats.items()}, "clocks": clocks,
        "final_norm_sqr": norm, "comm": comm, "jit": jit_info, "extras": extras,
    }
    print(json.dumps(line))
    if dist is not None:
        dist.barrier()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--qubits", type=int, default=30, help="qubits per GPU")
    ap.add_argument("--depth", type=int, default=40)
    ap.add_argument("--cpu-budget", type=float, default=20.0)
    ap.add_argument("--config5-local-qubits", type=int, default=33, help="N > 1: local qubits per GPU of the config-5 QFT (33 = 128 GiB per GPU)")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-extras", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--opt", action="append", default=[], metavar="NAME=VALUE",
                    help="engine option for an A/B run (qi_set_option), e.g. --opt lean=1; recorded in config.options")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
