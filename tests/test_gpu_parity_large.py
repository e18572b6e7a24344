"""Oracle parity at the sizes BASELINE.json names (`-m gpu`; the `large` marker: ~2-3 minutes of host time for the oracle).

tests/test_gpu_parity.py stops at 20 qubits for full-vector comparisons.  Here the GPU engine is compared with the oracle's
in-place arithmetic (oracle/qi_oracle.c `orc_gate`, the same per-gate arithmetic as the faithful pass, no clone) on states of
2^26 and 2^28 amplitudes -- where the 64-bit index paths, the phase tables over tile bits >= 16 (chunk tables 3 and 4) and the
CTA-tile scheduler's sliding tiles actually run:
  * config 2 (random layered H/RX/RZ/CNOT) at 26 and 28 qubits: norm, 64 probe amplitudes at splitmix-seeded indices and a
    3-term Pauli expectation value (SURVEY 8d's plan for this config);
  * a controlled-phase ladder circuit (QFT) on a GENERIC random state at 26 qubits: 64 probe amplitudes, so that every tile-chunk
    table is compared with the oracle and not only with the closed form on |+>;
  * config 3 at the full 24 sites: 2 first-order Trotter steps of the Heisenberg chain and <H>, against the oracle.
Every comparison runs on each executor of the product: the JIT tile modules, the interpreting tile kernel and the warp-window
kernel.  Tolerances are BASELINE.json's: 1e-12 absolute on amplitudes, 1e-10 relative on expectation values.
"""
import numpy as np
import pytest

from conftest import AMP_TOL, EXP_RTOL

pytestmark = [pytest.mark.gpu, pytest.mark.large]

EXECUTORS = [("tile_jit", {"tile": 1, "jit": 2}), ("tile_interpreter", {"tile": 1, "jit": 0}), ("window", {"tile": 0, "jit": 0})]


def _probe_indices(n, count=64, seed=20260007):
    from quant_iron_b200 import workloads as w
    rng = w.SplitMix64(seed)
    return [rng.next_u64() & ((1 << n) - 1) for _ in range(count)]


def _oracle_run(ref, v, n, specs):
    kinds = {"h": ref.G_H, "rx": ref.G_RX, "rz": ref.G_RZ, "cnot": ref.G_CNOT, "cp": ref.G_P, "swap": ref.G_SWAP}
    for name, targets, controls, params in specs:
        ref.gate_inplace(v, n, kinds[name], targets, controls, params)
    return v


def _with_options(gpu, opts):
    class Ctx:
        def __enter__(self):
            for k, val in opts.items():
                gpu.engine.set_option(k, val)

        def __exit__(self, *a):
            gpu.engine.set_option("tile", 1)
            gpu.engine.set_option("jit", 1)
    return Ctx()


def _three_terms(q, n):
    return q.SumOp([q.PauliString.new(0.7).with_op(0, q.Pauli.Z).with_op(1, q.Pauli.Z),
                    q.PauliString.new(-1.3).with_op(n // 2, q.Pauli.X),
                    q.PauliString.new(0.4).with_op(n - 1, q.Pauli.Y).with_op(3, q.Pauli.Z)])


@pytest.mark.parametrize("n,depth", [(26, 12), (28, 6)])
def test_config2_layered_circuit_probes_and_expectation_vs_oracle(gpu, ref, n, depth):
    from quant_iron_b200 import workloads as w
    specs = w.random_layered_circuit(n, depth)
    v = np.zeros(1 << n, dtype=np.complex128)
    v[0] = 1.0
    _oracle_run(ref, v, n, specs)
    idx = _probe_indices(n)
    want = v[idx]
    e_ref = _three_terms(ref, n).expectation_value(ref.State(v, n))
    assert np.max(np.abs(want)) > 0          # the probes see a generic state
    circuit = w.build_circuit(gpu, n, specs)
    for name, opts in EXECUTORS:
        with _with_options(gpu, opts):
            st = gpu.State.new_zero(n)
            gpu.engine.stats_reset()
            circuit.execute_(st)
            kernels = gpu.engine.stats()
            got = np.array([st.amplitude(i) for i in idx])
            err = float(np.max(np.abs(got - want)))
            assert err <= AMP_TOL, f"{name} n={n}: probe amplitude error {err:.3e} ({kernels})"
            assert abs(st.norm_sqr() - 1.0) <= 1e-10, name
            e_gpu = _three_terms(gpu, n).expectation_value(st)
            assert abs(e_gpu - e_ref) <= EXP_RTOL * max(1.0, abs(e_ref)), (name, e_gpu, e_ref)
            if name == "tile_jit":
                assert kernels.get("gate_tile_jit", {}).get("launches", 0) > 0 and "gate_window" not in kernels, kernels
            del st


def test_cp_ladders_on_a_generic_state_26q_vs_oracle(gpu, ref):
    """QFT (H + controlled-phase ladders + swaps) on a random state: every phase-table chunk (tile bits 0..20) against the oracle."""
    from quant_iron_b200 import workloads as w
    n = 26
    r0 = ref.random_state(n, 20260011)
    specs = w.qft_specs(n)
    circuit = w.build_circuit(gpu, n, specs)
    outs = {}
    for name, opts in EXECUTORS:
        with _with_options(gpu, opts):
            st = gpu.State(r0.state_vector, n)
            circuit.execute_(st)
            outs[name] = (np.array([st.amplitude(i) for i in _probe_indices(n)]), st.norm_sqr())
            del st
    v = r0.state_vector
    _oracle_run(ref, v, n, specs)
    want = v[_probe_indices(n)]
    for name, (got, nrm) in outs.items():
        err = float(np.max(np.abs(got - want)))
        assert err <= AMP_TOL, f"{name}: probe amplitude error {err:.3e}"
        assert abs(nrm - 1.0) <= 1e-10


def test_config3_heisenberg_24_sites_two_trotter_steps_vs_oracle(gpu, ref):
    n = 24
    args = (n, 1.0, 2.0, 3.0, 0.5, 0.1)
    h_ref, h_gpu = ref.heisenberg_1d(*args), gpu.heisenberg_1d(*args)
    s_ref = ref.trotter_evolve_state(h_ref, ref.State.new_plus(n), 0.01, 2, ref.TrotterOrder.First)
    e_ref = h_ref.expectation_value(s_ref)
    idx = _probe_indices(n)
    want = s_ref.state_vector[idx]
    for fuse in (1, 0):          # the fused 4-string window passes and the per-term kernels
        gpu.engine.set_option("fuse", fuse)
        try:
            s_gpu = gpu.trotter_evolve_state(h_gpu, gpu.State.new_plus(n), 0.01, 2, gpu.TrotterOrder.First)
            e_gpu = h_gpu.expectation_value(s_gpu)
        finally:
            gpu.engine.set_option("fuse", 1)
        got = np.array([s_gpu.amplitude(i) for i in idx])
        assert float(np.max(np.abs(got - want))) <= AMP_TOL, f"fuse={fuse}"
        assert abs(e_gpu - e_ref) <= EXP_RTOL * max(1.0, abs(e_ref)), (fuse, e_gpu, e_ref)
