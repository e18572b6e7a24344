"""qi_execute_host: Circuit::execute on a HOST-resident state vector with the PCIe copies overlapped
(csrc/host_pipeline.cu).

CPU tier: the split of the gate list into front / middle / back is a pure function of the list
(qi_host_pipeline_plan); executing the gates in that order on the ORACLE must give the state the original
order gives, and no front / back gate may touch a chunk-index qubit non-diagonally.
GPU tier: the pipelined entry against the oracle and against upload + execute_ + to_host.
"""
import ctypes as C

import numpy as np
import pytest

from conftest import AMP_TOL, vec

DIAG_KINDS = {4, 5, 6, 7, 8, 9, 10, 13}      # qi_gate_kind: Z, I, S, SDG, T, TDG, P, RZ act diagonally on their target


def _fuzz_builders(apis, n, seed, count=120, lazy_swaps=False):
    """The same random gate list through every API in `apis` (all operator kinds, 0-2 controls)."""
    rng = np.random.default_rng(seed)
    builders = [a.CircuitBuilder(n) for a in apis]
    for _ in range(count):
        kind = int(rng.integers(0, 16))
        qs = [int(q) for q in rng.permutation(n)[:4]]
        t, c1, c2, t2 = qs
        ang = float(rng.uniform(-3, 3))
        nc = int(rng.integers(0, 3))
        ctrls = [c1, c2][:nc]
        mt = int(rng.integers(0, n - 1))
        mc = [q for q in ctrls if q != mt]           # a control on the partner qubit mt + 1 is legal in the reference (operator.rs:940-1007)
        for b in builders:
            if kind == 0: b.ch_gates([t], ctrls) if nc else b.h_gate(t)
            elif kind == 1: b.cx_gates([t], ctrls) if nc else b.x_gate(t)
            elif kind == 2: b.cy_gates([t], ctrls) if nc else b.y_gate(t)
            elif kind == 3: b.cz_gates([t], ctrls) if nc else b.z_gate(t)
            elif kind == 4: b.cs_gates([t], ctrls) if nc else b.t_gate(t)
            elif kind == 5: b.cp_gates([t], ctrls, ang) if nc else b.p_gate(t, ang)
            elif kind == 6: b.crx_gates([t], ctrls, ang) if nc else b.rx_gate(t, ang)
            elif kind == 7: b.cry_gates([t], ctrls, ang) if nc else b.ry_gate(t, ang)
            elif kind == 8: b.crz_gates([t], ctrls, ang) if nc else b.rz_gate(t, ang)
            elif kind == 9: b.swap_gate(t, t2) if lazy_swaps else b.cswap_gate(t, t2, [c1])
            elif kind == 10: b.cswap_gate(t, t2, [c1])
            elif kind == 11: b.toffoli_gate(c1, c2, t)
            elif kind == 12: b.ry_phase_gate(t, ang, 0.5 * ang)
            elif kind == 13: b.cmatchgate(mt, mc, ang, 0.3 * ang, -0.7 * ang) if mc else b.matchgate(mt, ang, 0.3 * ang, -0.7 * ang)
            elif kind == 14: b.rz_gate(t, ang)
            else: b.cnot_gate(t, c1)
    return [b.build() for b in builders]


def _plan(circuit, n, k):
    from quant_iron_b200 import _ffi
    runs = circuit._lower()
    assert len(runs) == 1 and runs[0][0] == "ops"
    arr, count = runs[0][1], runs[0][2]
    order = (C.c_uint64 * max(1, count))()
    nf, nm, nb = C.c_uint64(), C.c_uint64(), C.c_uint64()
    _ffi.check(_ffi.lib.qi_host_pipeline_plan(n, arr, count, k, order, C.byref(nf), C.byref(nm), C.byref(nb)))
    assert nf.value + nm.value + nb.value == count
    return [int(order[i]) for i in range(count)], nf.value, nm.value, nb.value


def _nondiag_targets(rec):
    """qubits a C-ABI gate record uses non-diagonally"""
    if rec.kind in DIAG_KINDS:
        return set()
    t = {int(rec.targets[j]) for j in range(rec.num_targets)}
    if rec.kind == 18:                          # Matchgate acts on (q, q + 1)
        partner = int(rec.targets[0]) + 1
        if any(int(rec.controls[j]) == partner for j in range(rec.num_controls)):
            return set()                        # control on the partner: only the phase on |11> survives (a diagonal gate)
        t.add(partner)
    return t


@pytest.mark.parametrize("n,k,seed", [(8, 1, 1), (9, 2, 2), (10, 3, 3), (10, 3, 4), (11, 4, 5), (9, 3, 6), (10, 2, 7)])
def test_plan_is_a_valid_reordering(ref, n, k, seed):
    import quant_iron_b200 as gpu       # host-side lowering only: no device call
    cg, cr = _fuzz_builders([gpu, ref], n, seed, lazy_swaps=(seed % 2 == 0))
    assert len(cg.gates) == len(cr.gates)
    order, nf, nm, nb = _plan(cg, n, k)
    assert sorted(order) == list(range(len(cr.gates)))
    for sect in (order[:nf], order[nf:nf + nm], order[nf + nm:]):
        assert sect == sorted(sect)                                   # circuit order inside every section
    top = set(range(n - k, n))
    recs = cg._lower()[0][1]
    for i in order[:nf] + order[nf + nm:]:
        assert not (_nondiag_targets(recs[i]) & top), (i, recs[i].kind)
    start = ref.random_state(n, 900 + seed)
    want = cr.execute(start)
    got = ref.Circuit.with_gates([cr.gates[i] for i in order], n).execute(start)
    assert float(np.max(np.abs(vec(got) - vec(want)))) <= 1e-13


def test_plan_of_the_benchmark_circuit():
    """30 qubits, depth 40, 8 chunks: the two light-cone-free trapezoids hold more than half of the gates, so more than
    half of the circuit overlaps the 2 x 16 GiB of PCIe traffic."""
    import quant_iron_b200 as gpu
    from quant_iron_b200 import workloads as w
    specs = w.random_layered_circuit(30, 40)
    order, nf, nm, nb = _plan(w.build_circuit(gpu, 30, specs), 30, 3)
    assert nf + nm + nb == 1780
    assert nf >= 450 and nb >= 450, (nf, nm, nb)
    # no chunking: everything is "front"
    _, nf0, nm0, nb0 = _plan(w.build_circuit(gpu, 30, specs), 30, 0)
    assert (nf0, nm0, nb0) == (1780, 0, 0)


@pytest.fixture
def small_pipeline(gpu):
    gpu.engine.set_option("host_min_qubits", 0)
    yield
    gpu.engine.set_option("host_min_qubits", 26)
    gpu.engine.set_option("host_chunk_qubits", 3)


@pytest.mark.gpu
@pytest.mark.unproven
@pytest.mark.parametrize("n,k,seed", [(11, 1, 1), (11, 3, 2), (12, 2, 3), (12, 3, 4), (13, 3, 5), (13, 4, 6), (14, 3, 7)])
def test_execute_host_pipelined_vs_oracle(gpu, ref, small_pipeline, n, k, seed):
    gpu.engine.set_option("host_chunk_qubits", k)
    cg, cr = _fuzz_builders([gpu, ref], n, seed, count=160)
    start = ref.random_state(n, 700 + seed)
    want = vec(cr.execute(start))
    dev = gpu.State.new_zero(n)
    buf = np.array(start.state_vector, dtype=np.complex128)
    out = cg.execute_host_(dev, buf)
    assert float(np.max(np.abs(out - want))) <= AMP_TOL
    assert float(np.max(np.abs(vec(dev) - want))) <= AMP_TOL           # the working buffer holds the final state
    cg.execute_host_(dev, buf, buf)                                     # in place on the host
    assert float(np.max(np.abs(buf - want))) <= AMP_TOL


@pytest.mark.gpu
@pytest.mark.unproven
@pytest.mark.parametrize("n", [11, 13])
def test_execute_host_plain_sequence_paths(gpu, ref, small_pipeline, n):
    """Lists with relabelled SWAPs (the QFT) and k = 0 take upload / execute / download: same answers."""
    from quant_iron_b200 import workloads as w
    specs = w.random_layered_circuit(n, 5) + w.qft_specs(n)
    cg, cr = w.build_circuit(gpu, n, specs), w.build_circuit(ref, n, specs)
    start = ref.random_state(n, 31)
    want = vec(cr.execute(start))
    dev = gpu.State.new_zero(n)
    out = cg.execute_host_(dev, np.array(start.state_vector))
    assert float(np.max(np.abs(out - want))) <= AMP_TOL
    gpu.engine.set_option("host_chunk_qubits", 0)
    cl, clr = _fuzz_builders([gpu, ref], n, 11)
    out = cl.execute_host_(dev, np.array(start.state_vector))
    assert float(np.max(np.abs(out - vec(clr.execute(start))))) <= AMP_TOL


@pytest.mark.gpu
@pytest.mark.unproven
@pytest.mark.parametrize("n,depth", [(16, 12), (20, 40)])
def test_execute_host_layered_circuit_vs_resident_path(gpu, ref, small_pipeline, n, depth):
    from quant_iron_b200 import workloads as w
    specs = w.random_layered_circuit(n, depth)
    cg = w.build_circuit(gpu, n, specs)
    start = np.array(ref.random_state(n, 5).state_vector)
    dev = gpu.State.new_zero(n)
    out = cg.execute_host_(dev, start)
    resident = gpu.State(start, n)
    cg.execute_(resident)
    assert float(np.max(np.abs(out - vec(resident)))) <= 1e-13
    if n <= 16:
        want = vec(w.build_circuit(ref, n, specs).execute(ref.State(start, n)))
        assert float(np.max(np.abs(out - want))) <= AMP_TOL


@pytest.mark.gpu
@pytest.mark.unproven
def test_execute_host_default_options_26_qubits(gpu):
    """Default options: a 26-qubit state (1 GiB) is pipelined in 8 chunks of 128 MiB; the result equals the plain
    upload / execute / download sequence on the same device buffer."""
    from quant_iron_b200 import workloads as w
    n = 26
    specs = w.random_layered_circuit(n, 12)
    cg = w.build_circuit(gpu, n, specs)
    start = np.zeros(1 << n, dtype=np.complex128)
    start[0] = 1.0
    dev = gpu.State.new_zero(n)
    out = cg.execute_host_(dev, start)
    dev.upload_(start)
    cg.execute_(dev)
    plain = dev.to_host()
    assert float(np.max(np.abs(out - plain))) <= 1e-13
    assert abs(float(np.vdot(out, out).real) - 1.0) <= 1e-10
