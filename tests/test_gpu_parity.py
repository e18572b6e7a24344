"""GPU engine vs CPU oracle on the same seeded inputs (the parity tests proper, `-m gpu`).

Bars (BASELINE.json north star): max abs amplitude error <= 1e-12; expectation values within 1e-10
relative; seeded measurement bins bit-identical.  Sizes are chosen so the oracle finishes in seconds;
the BASELINE.json full sizes are covered by size-independent properties at the bottom.
"""
import math

import numpy as np
import pytest

from conftest import AMP_TOL, EXP_RTOL, assert_amps, vec

pytestmark = pytest.mark.gpu

PI = math.pi


@pytest.fixture(autouse=True, params=[18, 11], ids=["tile_from_18q", "tile_from_11q"])
def _tile_threshold(request):
    """Every test runs twice: with the default split between the warp-tile kernel (k_window, < 18 local qubits) and the
    CTA-tile kernel (k_tile), and with k_tile taking every state it can hold (>= 11 qubits), so that the small oracle-sized
    cases exercise the kernel the benchmark sizes run on."""
    import quant_iron_b200
    quant_iron_b200.engine.set_option("tile_min_qubits", request.param)
    quant_iron_b200.engine.set_option("tile_min_gates", 1 if request.param == 11 else 3)      # 11: lone gates take the tile kernel too
    yield
    quant_iron_b200.engine.set_option("tile_min_qubits", 18)
    quant_iron_b200.engine.set_option("tile_min_gates", 3)


def _pair(gpu, ref, n, seed=20260002):
    r = ref.random_state(n, seed)
    return gpu.State(r.state_vector, n), r


GATES_1Q = [("h", ()), ("x", ()), ("y", ()), ("z", ()), ("s", ()), ("t", ()), ("s_dag", ()), ("t_dag", ()),
            ("p", (0.37,)), ("rx", (1.1,)), ("ry", (-0.7,)), ("rz", (2.3,))]


@pytest.mark.parametrize("path", [1, 0])
@pytest.mark.parametrize("n", [1, 2, 3, 5, 9, 11, 14])
def test_every_single_qubit_gate_every_target(gpu, ref, n, path):
    """each gate on every target qubit, 0/1/2 controls, random normalised state."""
    gpu.engine.set_option("path", path)
    try:
        g, r = _pair(gpu, ref, n)
        targets = range(n) if n <= 11 else [0, 1, 4, 5, 6, n - 2, n - 1]
        for name, args in GATES_1Q:
            for t in targets:
                g = getattr(g, name)(t, *args)
                r = getattr(r, name)(t, *args)
            assert_amps(g, vec(r), msg=f"{name} n={n}")
            others = [q for q in range(n)]
            if n >= 2:
                t, c = others[0], others[-1]
                g = getattr(g, f"c{name}_multi")([t], [c], *args)
                r = getattr(r, f"c{name}_multi")([t], [c], *args)
                g = getattr(g, f"c{name}_multi")([c], [t], *args)
                r = getattr(r, f"c{name}_multi")([c], [t], *args)
            if n >= 3:
                t, c = n // 2, [0, n - 1]
                g = getattr(g, f"c{name}_multi")([t], c, *args)
                r = getattr(r, f"c{name}_multi")([t], c, *args)
            assert_amps(g, vec(r), msg=f"c{name} n={n}")
    finally:
        gpu.engine.set_option("path", 0)


@pytest.mark.parametrize("path", [1, 0])
@pytest.mark.parametrize("n", [2, 3, 6, 11, 13])
def test_two_qubit_operators(gpu, ref, n, path):
    gpu.engine.set_option("path", path)
    try:
        g, r = _pair(gpu, ref, n, seed=99)
        u = [[0.6 + 0j, 0.8j], [0.8j, 0.6 + 0j]]
        for a in range(n):
            for b in range(n):
                if a == b:
                    continue
                g, r = g.swap(a, b), r.swap(a, b)
                g, r = g.cnot(a, b), r.cnot(a, b)
                if n <= 6 or (a + b) % 3 == 0:
                    g, r = g.cunitary_multi([a], [b], u), r.cunitary_multi([a], [b], u)
            if a + 1 < n:
                g, r = g.matchgate(a, 0.7, 0.4, 1.1), r.matchgate(a, 0.7, 0.4, 1.1)
        assert_amps(g, vec(r), msg=f"swap/cnot/matchgate n={n}")
        if n >= 3:
            g, r = g.toffoli(0, 1, 2), r.toffoli(0, 1, 2)
            g, r = g.cswap(0, n - 1, [1]), r.cswap(0, n - 1, [1])
            g, r = g.cmatchgate(0, 1.3, 0.2, -0.5, [n - 1]), r.cmatchgate(0, 1.3, 0.2, -0.5, [n - 1])
            g, r = g.ry_phase(1, 0.9, 0.3), r.ry_phase(1, 0.9, 0.3)
            assert_amps(g, vec(r), msg=f"3-qubit ops n={n}")
    finally:
        gpu.engine.set_option("path", 0)


@pytest.mark.parametrize("path", [1, 0])
@pytest.mark.parametrize("n,depth", [(4, 8), (9, 8), (12, 10), (16, 12), (20, 40)])
def test_random_layered_circuit_matches_oracle(gpu, ref, n, depth, path):
    """BASELINE config 2's generator at oracle-sized n: full-amplitude comparison."""
    from quant_iron_b200 import workloads as w
    gpu.engine.set_option("path", path)
    try:
        specs = w.random_layered_circuit(n, depth)
        out_g = w.build_circuit(gpu, n, specs).execute(gpu.State.new_zero(n))
        out_r = w.build_circuit(ref, n, specs).execute(ref.State.new_zero(n))
        assert_amps(out_g, vec(out_r), msg=f"layered n={n} depth={depth}")
        assert abs(out_g.norm_sqr() - 1.0) < 1e-12
    finally:
        gpu.engine.set_option("path", 0)


@pytest.mark.parametrize("n", [3, 8, 11, 16, 20])
def test_qft_matches_oracle_and_closed_forms(gpu, ref, n):
    """BASELINE config 1 (n=20): Subroutine::qft via CircuitBuilder on new_plus -> |0...0>; a basis
    state -> the both-indices-bit-reversed DFT; iqft . qft = identity (unpinned by the
    reference's tests, pinned here by closed forms and by the oracle)."""
    qs = list(range(n))
    cg = gpu.CircuitBuilder(n).add_subroutine(gpu.Subroutine.qft(qs, n)).build()
    cr = ref.CircuitBuilder(n).add_subroutine(ref.Subroutine.qft(qs, n)).build()
    out_g = cg.execute(gpu.State.new_plus(n))
    out_r = cr.execute(ref.State.new_plus(n))
    assert_amps(out_g, vec(out_r), msg=f"qft|+> n={n}")
    v = vec(out_g)
    assert abs(v[0] - 1.0) <= 1e-12 and np.max(np.abs(v[1:])) <= 1e-12
    if n <= 11:
        x = (0b1011 % (1 << n)) | 1
        out = vec(cg.execute(gpu.State.new_basis_n(n, x)))
        N = 1 << n

        def rev(k):
            return int(format(k, f"0{n}b")[::-1], 2)
        exp = np.array([np.exp(2j * PI * rev(x) * rev(k) / N) for k in range(N)]) / math.sqrt(N)
        assert np.max(np.abs(out - exp)) <= 1e-12
    g, r = _pair(gpu, ref, n, seed=5)
    ci = gpu.CircuitBuilder(n).add_subroutine(gpu.Subroutine.qft(qs, n)).add_subroutine(gpu.Subroutine.iqft(qs, n)).build()
    assert_amps(ci.execute(g), vec(r), msg="iqft.qft")


@pytest.mark.parametrize("n", [2, 6, 10, 14])
def test_pauli_strings_vs_oracle(gpu, ref, n):
    g, r = _pair(gpu, ref, n, seed=7)
    rng = np.random.default_rng(n)
    for trial in range(12):
        k = int(rng.integers(0, min(n, 5) + 1))
        qs = [int(q) for q in rng.choice(n, size=k, replace=False)]
        ps_g, ps_r = gpu.PauliString.new(complex(0.3, -0.2)), ref.PauliString.new(complex(0.3, -0.2))
        for q in qs:
            which = int(rng.integers(0, 3))
            ps_g.add_op(q, [gpu.Pauli.X, gpu.Pauli.Y, gpu.Pauli.Z][which])
            ps_r.add_op(q, [ref.Pauli.X, ref.Pauli.Y, ref.Pauli.Z][which])
        assert_amps(ps_g.apply(g), vec(ps_r.apply(r)), msg="apply")
        assert_amps(ps_g.apply_normalised(g), vec(ps_r.apply_normalised(r)), msg="apply_normalised")
        assert_amps(ps_g.apply_exp(g), vec(ps_r.apply_exp(r)), msg="apply_exp")
        f = complex(0.1 * trial, -0.25)
        g, r = ps_g.apply_exp_factor(g, f), ps_r.apply_exp_factor(r, f)
        nrm = math.sqrt(r.inner_product(r).real)
        assert_amps(g, vec(r), tol=AMP_TOL * max(1.0, nrm), msg="apply_exp_factor chain")
        g, r = g.normalise(), r.normalise()


@pytest.mark.parametrize("n,steps", [(4, 5), (10, 5), (16, 3)])
def test_heisenberg_trotter_expectation_vs_oracle(gpu, ref, n, steps):
    """BASELINE config 3 at oracle-sized n: heisenberg_1d(n,1,2,3,0.5,0.1), new_plus, first- and
    second-order Trotter, SumOp::expectation_value."""
    hg, hr = gpu.heisenberg_1d(n, 1.0, 2.0, 3.0, 0.5, 0.1), ref.heisenberg_1d(n, 1.0, 2.0, 3.0, 0.5, 0.1)
    assert hg.num_terms() == hr.num_terms() == 4 * n
    for order_g, order_r in ((gpu.TrotterOrder.First, ref.TrotterOrder.First), (gpu.TrotterOrder.Second, ref.TrotterOrder.Second)):
        sg = gpu.trotter_evolve_state(hg, gpu.State.new_plus(n), 0.01, steps, order_g)
        sr = ref.trotter_evolve_state(hr, ref.State.new_plus(n), 0.01, steps, order_r)
        assert_amps(sg, vec(sr), msg="trotter state")
        eg, er = hg.expectation_value(sg), hr.expectation_value(sr)
        assert abs(eg - er) <= EXP_RTOL * abs(er), (eg, er)
    ag, ar = hg.apply(sg), hr.apply(sr)
    assert_amps(ag, vec(ar), tol=1e-11, msg="SumOp.apply")   # |H psi| ~ 10: bar scaled by the amplitude size


@pytest.mark.parametrize("n", [1, 4, 10, 15])
def test_linear_algebra_vs_oracle(gpu, ref, n):
    g1, r1 = _pair(gpu, ref, n, seed=1)
    g2, r2 = _pair(gpu, ref, n, seed=2)
    ig, ir = g1.inner_product(g2), r1.inner_product(r2)
    assert abs(ig - ir) <= 1e-12
    assert abs(g1.norm_sqr() - 1.0) <= 1e-12
    assert_amps(g1 + g2, vec(r1 + r2))
    assert_amps(g1 - g2, vec(r1 - r2))
    assert_amps(g1 * complex(0.3, 0.4), vec(r1 * complex(0.3, 0.4)))
    assert_amps((g1 + g2).normalise(), vec((r1 + r2).normalise()))
    assert_amps(g1.conj(), vec(r1.conj()))
    if n <= 7:
        assert_amps(g1.tensor_product(g2), vec(r1.tensor_product(r2)))


@pytest.mark.parametrize("n,qubits", [(3, [0]), (3, []), (8, [1, 6]), (12, [0, 3, 5]), (12, [11, 2, 7, 0]),
                                      (12, []), (14, [13]), (16, [4, 5, 6, 7, 8, 9, 10, 11, 12, 13])])
def test_probabilities_and_seeded_sampling_bit_identical(gpu, ref, n, qubits):
    """Shared-seed contract: the bins drawn on the device equal the oracle's, draw for draw."""
    g, r = _pair(gpu, ref, n, seed=11)
    pg, pr = g.probabilities(qubits), r.probabilities(qubits if qubits else list(range(n)))
    assert np.max(np.abs(pg - pr)) <= 1e-13
    shots, seed = 512, 20260003
    assert r.sample_margin(qubits, shots, seed) > 1e-11, "a draw sits on a CDF edge: pick another seed"
    assert np.array_equal(g.sample_counts(qubits, shots, seed), r.sample_counts(qubits, shots, seed))


@pytest.mark.parametrize("n", [2, 5, 10])
def test_measure_all_bases_same_outcomes_and_states_as_oracle(gpu, ref, n):
    g, r = _pair(gpu, ref, n, seed=13)
    u = [[0.6 + 0j, 0.8j], [0.8j, 0.6 + 0j]]
    bases = [(gpu.MeasurementBasis.Computational, ref.MeasurementBasis.Computational),
             (gpu.MeasurementBasis.X, ref.MeasurementBasis.X), (gpu.MeasurementBasis.Y, ref.MeasurementBasis.Y),
             (gpu.MeasurementBasis.Custom(u), ref.MeasurementBasis.Custom(u))]
    for bg, br in bases:
        for qubits in ([0], [n - 1, 0], []):
            for seed in (1, 2, 3):
                mg, mr = g.measure(bg, qubits, seed=seed), r.measure(br, qubits, seed=seed)
                assert mg.get_outcomes() == mr.get_outcomes()
                assert mg.get_indices() == mr.get_indices()
                assert_amps(mg.get_new_state(), vec(mr.get_new_state()), msg=f"collapsed {bg} {qubits}")
    rs_g = g.measure_n(gpu.MeasurementBasis.Computational, [0, 1], 6, seed=42)
    rs_r = r.measure_n(ref.MeasurementBasis.Computational, [0, 1], 6, seed=42)
    assert [m.get_outcomes() for m in rs_g] == [m.get_outcomes() for m in rs_r]


def test_circuit_with_measurement_and_pauli_gates(gpu, ref):
    n = 6
    ps = lambda q: q.PauliString.new(0.8).with_op(1, q.Pauli.X).with_op(4, q.Pauli.Z)  # noqa: E731

    def build(q):
        return (q.CircuitBuilder(n).h_gates(list(range(n))).cnot_gate(1, 0).rz_gate(2, 0.4)
                .pauli_time_evolution_gate(ps(q), 0.3).measure_gate(q.MeasurementBasis.X, [2, 3])
                .pauli_string_gate(ps(q)).toffoli_gate(0, 1, 5).build())
    out_g = build(gpu).execute(gpu.State.new_zero(n), seed=77)
    out_r = build(ref).execute(ref.State.new_zero(n), seed=77)
    assert_amps(out_g, vec(out_r))


# ---- BASELINE.json full sizes: size-independent properties (no oracle at this size) -----------------
@pytest.mark.parametrize("n", [26, 28])
def test_large_state_properties(gpu, n):
    """norm preservation, involutions, and the QFT closed form at sizes the oracle cannot hold."""
    st = gpu.State.new_random(n)
    assert abs(st.norm_sqr() - 1.0) < 1e-10
    probe = [st.amplitude(i) for i in (0, 12345, (1 << n) - 1)]
    for t in (0, 3, 7, n // 2, n - 1):
        st.h_(t).rx_(t, 0.3).rz_(t, 1.1).rz_(t, -1.1).rx_(t, -0.3).h_(t)
    st.cnot_(n - 1, 0).cnot_(n - 1, 0).swap_(1, n - 2).swap_(1, n - 2)
    assert abs(st.norm_sqr() - 1.0) < 1e-10
    for i, a in zip((0, 12345, (1 << n) - 1), probe):
        assert abs(st.amplitude(i) - a) < 1e-12
    del st
    plus = gpu.State.new_plus(n)
    qs = list(range(n))
    gpu.CircuitBuilder(n).add_subroutine(gpu.Subroutine.qft(qs, n)).build().execute_(plus)
    assert abs(plus.amplitude(0) - 1.0) < 1e-12
    assert abs(plus.norm_sqr() - 1.0) < 1e-10
    assert abs(plus.amplitude(1)) < 1e-12 and abs(plus.amplitude((1 << n) - 1)) < 1e-12


@pytest.mark.parametrize("n,seed", [(9, 1), (10, 2), (11, 3), (12, 4), (13, 5), (12, 6)])
def test_fuzz_random_gate_lists_through_fused_executor(gpu, ref, n, seed):
    """Random gate lists over every operator kind with random controls, executed as ONE circuit (fused
    window passes, merged phase tables, lazy SWAP relabelling) vs gate-by-gate on the oracle."""
    rng = np.random.default_rng(seed)
    bg, br = gpu.CircuitBuilder(n), ref.CircuitBuilder(n)
    for _ in range(160):
        kind = int(rng.integers(0, 14))
        qs = [int(q) for q in rng.permutation(n)[:4]]
        t, c1, c2, t2 = qs
        ang = float(rng.uniform(-3, 3))
        nc = int(rng.integers(0, 3))
        ctrls = [c1, c2][:nc]
        for b in (bg, br):
            if kind == 0: b.ch_gates([t], ctrls) if nc else b.h_gate(t)
            elif kind == 1: b.cx_gates([t], ctrls) if nc else b.x_gate(t)
            elif kind == 2: b.cy_gates([t], ctrls) if nc else b.y_gate(t)
            elif kind == 3: b.cz_gates([t], ctrls) if nc else b.z_gate(t)
            elif kind == 4: b.cs_gates([t], ctrls) if nc else b.t_gate(t)
            elif kind == 5: b.cp_gates([t], ctrls, ang) if nc else b.p_gate(t, ang)
            elif kind == 6: b.crx_gates([t], ctrls, ang) if nc else b.rx_gate(t, ang)
            elif kind == 7: b.cry_gates([t], ctrls, ang) if nc else b.ry_gate(t, ang)
            elif kind == 8: b.crz_gates([t], ctrls, ang) if nc else b.rz_gate(t, ang)
            elif kind == 9: b.swap_gate(t, t2)
            elif kind == 10: b.cswap_gate(t, t2, [c1])
            elif kind == 11: b.toffoli_gate(c1, c2, t)
            elif kind == 12: b.ry_phase_gate(t, ang, 0.5 * ang)
            else: b.cnot_gate(t, c1)
    sg, sr = _pair(gpu, ref, n, seed=100 + seed)
    out_g = bg.build().execute(sg)
    out_r = br.build().execute(sr)
    assert_amps(out_g, vec(out_r), msg=f"fuzz n={n} seed={seed}")
    # the lazily relabelled layout must be invisible to every read path
    for i in (0, 1, (1 << n) - 1, 37 % (1 << n)):
        assert abs(out_g.amplitude(i) - complex(out_r.state_vector[i])) <= AMP_TOL
    qs = [0, n - 1, n // 2]
    assert np.max(np.abs(out_g.probabilities(qs) - out_r.probabilities(qs))) <= 1e-13
    assert abs(out_g.inner_product(sg) - out_r.inner_product(sr)) <= 1e-12


@pytest.mark.parametrize("n,seed", [(9, 0), (10, 1), (12, 2), (14, 3)])
def test_pauli_exp_sequence_fused_passes_vs_oracle(gpu, ref, n, seed):
    """qi_apply_pauli_exp_sequence: random strings (X/Y/Z anywhere, up to 7 factors, complex
    factors, an empty string) applied as ONE fused sequence vs one apply_exp_factor call per string on
    the oracle, and vs the engine's own per-term kernels (fuse = 0)."""
    from quant_iron_b200.pauli import apply_exp_sequence_
    rng = np.random.default_rng(seed)
    g, r = _pair(gpu, ref, n, seed=300 + seed)
    strings_g, strings_r, factors = [], [], []
    for trial in range(60):
        k = int(rng.integers(0, 8)) if trial != 7 else 0
        qs = [int(q) for q in rng.choice(n, size=min(k, n), replace=False)]
        c = complex(rng.uniform(-1, 1), rng.uniform(-0.3, 0.3) if trial % 3 == 0 else 0.0)
        pg, pr = gpu.PauliString.new(c), ref.PauliString.new(c)
        for q in qs:
            which = int(rng.integers(0, 3))
            pg.add_op(q, [gpu.Pauli.X, gpu.Pauli.Y, gpu.Pauli.Z][which])
            pr.add_op(q, [ref.Pauli.X, ref.Pauli.Y, ref.Pauli.Z][which])
        strings_g.append(pg)
        strings_r.append(pr)
        factors.append(complex(rng.uniform(-0.2, 0.2), rng.uniform(-0.5, 0.5)))
    for pr, f in zip(strings_r, factors):
        r = pr.apply_exp_factor(r, f)
    fused = apply_exp_sequence_(g.clone(), strings_g, factors)
    nrm = max(1.0, math.sqrt(r.inner_product(r).real))
    assert_amps(fused, vec(r), tol=AMP_TOL * nrm, msg="fused sequence vs oracle")
    gpu.engine.set_option("fuse", 0)
    try:
        unfused = apply_exp_sequence_(g.clone(), strings_g, factors)
    finally:
        gpu.engine.set_option("fuse", 1)
    assert_amps(unfused, vec(r), tol=AMP_TOL * nrm, msg="per-term sequence vs oracle")
    assert np.max(np.abs(vec(fused) - vec(unfused))) <= 1e-14 * nrm


def test_pauli_time_evolution_gates_in_a_circuit(gpu, ref):
    """Gate::PauliTimeEvolution runs (gate.rs:116-118) inside a circuit between operator gates."""
    n = 11
    bg, br = gpu.CircuitBuilder(n), ref.CircuitBuilder(n)
    hg, hr = gpu.heisenberg_1d(n, 1.0, 0.5, -0.7, 0.3, 0.2), ref.heisenberg_1d(n, 1.0, 0.5, -0.7, 0.3, 0.2)
    for b, h in ((bg, hg), (br, hr)):
        b.h_gates(list(range(n)))
        for t in h.terms:
            b.pauli_time_evolution_gate(t, 0.05)
        b.cnot_gate(3, 9)
        for t in h.terms[::-1]:
            b.pauli_time_evolution_gate(t, 0.02)
        b.rx_gate(10, 0.4)
    out_g = bg.build().execute(gpu.State.new_zero(n))
    out_r = br.build().execute(ref.State.new_zero(n))
    assert_amps(out_g, vec(out_r), msg="circuit with PauliTimeEvolution runs")
    bad = gpu.CircuitBuilder(n).pauli_time_evolution_gate(gpu.PauliString.new(1j).with_op(0, gpu.Pauli.X), 0.1).build()
    with pytest.raises(gpu.Error) as e:
        bad.execute(gpu.State.new_zero(n))
    assert e.value.variant == "InvalidPauliStringCoefficient"


def test_trotter_24_sites_fused_vs_per_term(gpu):
    """BASELINE config 3's size: the fused Trotter path and the per-term path agree to rounding, the norm
    is preserved, and the fused path really ran window passes (kernel statistics)."""
    n = 24
    h = gpu.heisenberg_1d(n, 1.0, 2.0, 3.0, 0.5, 0.1)
    gpu.engine.stats_reset()
    a = gpu.trotter_evolve_state(h, gpu.State.new_plus(n), 0.01, 3, gpu.TrotterOrder.Second)
    stats = {k: v["launches"] for k, v in gpu.engine.stats().items()}
    assert stats.get("pauli_exp_window", 0) > 0 and stats.get("pauli_exp", 0) == 0, stats
    gpu.engine.set_option("fuse", 0)
    try:
        b = gpu.trotter_evolve_state(h, gpu.State.new_plus(n), 0.01, 3, gpu.TrotterOrder.Second)
    finally:
        gpu.engine.set_option("fuse", 1)
    assert abs(a.norm_sqr() - 1.0) < 1e-12
    d = a - b
    assert d.norm_sqr() < 1e-24
    ea, eb = h.expectation_value(a), h.expectation_value(b)
    assert abs(ea - eb) <= EXP_RTOL * abs(eb)


@pytest.mark.parametrize("n,regs", [(9, 4), (13, 4), (16, 4), (12, 3), (14, 5)])
def test_tma_prefetched_window_kernel_matches_oracle(gpu, ref, n, regs):
    """k_window_tma (cp.async.bulk + mbarrier staging, option "tma" = 1) computes exactly what the direct
    kernel computes: layered circuit and QFT vs the oracle, for every register-window width."""
    from quant_iron_b200 import workloads as w
    gpu.engine.set_option("tma", 1)
    gpu.engine.set_option("window_regs", regs)
    try:
        specs = w.random_layered_circuit(n, 10)
        out_g = w.build_circuit(gpu, n, specs).execute(gpu.State.new_zero(n))
        out_r = w.build_circuit(ref, n, specs).execute(ref.State.new_zero(n))
        assert_amps(out_g, vec(out_r), msg=f"tma layered n={n}")
        qs = list(range(n))
        cg = gpu.CircuitBuilder(n).add_subroutine(gpu.Subroutine.qft(qs, n)).build()
        cr = ref.CircuitBuilder(n).add_subroutine(ref.Subroutine.qft(qs, n)).build()
        g, r = _pair(gpu, ref, n, seed=77)
        assert_amps(cg.execute(g), vec(cr.execute(r)), msg=f"tma qft n={n}")
    finally:
        gpu.engine.set_option("tma", 0)
        gpu.engine.set_option("window_regs", 0)


@pytest.mark.parametrize("n,seed", [(9, 0), (11, 1), (14, 2)])
def test_sumop_expectation_batched_vs_oracle(gpu, ref, n, seed):
    """SumOp::expectation_value with many random terms (identity term, strings wider than a register window,
    complex coefficients): the batched read-only window passes, the per-term kernels (fuse = 0) and the
    oracle agree within the 1e-10 relative bar."""
    rng = np.random.default_rng(50 + seed)
    g, r = _pair(gpu, ref, n, seed=400 + seed)
    tg, tr = [], []
    for trial in range(90):
        k = 0 if trial == 5 else int(rng.integers(1, min(n, 8) + 1))
        qs = [int(q) for q in rng.choice(n, size=k, replace=False)]
        c = complex(rng.uniform(-1, 1), rng.uniform(-1, 1) if trial % 4 == 0 else 0.0)
        pg, pr = gpu.PauliString.new(c), ref.PauliString.new(c)
        for q in qs:
            which = int(rng.integers(0, 3))
            pg.add_op(q, [gpu.Pauli.X, gpu.Pauli.Y, gpu.Pauli.Z][which])
            pr.add_op(q, [ref.Pauli.X, ref.Pauli.Y, ref.Pauli.Z][which])
        tg.append(pg)
        tr.append(pr)
    er = ref.SumOp(tr).expectation_value(r)
    gpu.engine.stats_reset()
    eg = gpu.SumOp(tg).expectation_value(g)
    launches = gpu.engine.stats()["pauli_expect"]["launches"]
    assert launches < 60, launches          # 90 terms share far fewer passes
    gpu.engine.set_option("fuse", 0)
    try:
        eu = gpu.SumOp(tg).expectation_value(g)
    finally:
        gpu.engine.set_option("fuse", 1)
    assert abs(eg - er) <= EXP_RTOL * abs(er), (eg, er)
    assert abs(eu - er) <= EXP_RTOL * abs(er), (eu, er)
    # per-term check: every single string alone
    for pg, pr in list(zip(tg, tr))[:20]:
        a, b = gpu.SumOp([pg]).expectation_value(g), ref.SumOp([pr]).expectation_value(r)
        assert abs(a - b) <= 1e-12, (a, b)


def test_pauli_sequences_longer_than_one_launch(gpu, ref):
    """More terms than one window launch holds (192 ops) and more expectation terms than one group holds: the
    scheduler must split, never truncate."""
    from quant_iron_b200.pauli import apply_exp_sequence_
    n = 10
    rng = np.random.default_rng(9)
    g, r = _pair(gpu, ref, n, seed=500)
    sg, sr = [], []
    for i in range(700):
        q = int(rng.integers(0, n))
        which = int(rng.integers(0, 3))
        c = complex(rng.uniform(-1, 1), 0.0)
        sg.append(gpu.PauliString.new(c).with_op(q, [gpu.Pauli.X, gpu.Pauli.Y, gpu.Pauli.Z][which]))
        sr.append(ref.PauliString.new(c).with_op(q, [ref.Pauli.X, ref.Pauli.Y, ref.Pauli.Z][which]))
    gpu.engine.stats_reset()
    out = apply_exp_sequence_(g.clone(), sg, [complex(0.0, -0.01)] * len(sg))
    launches = gpu.engine.stats()["pauli_exp_window"]["launches"]
    assert 4 <= launches <= 40, launches
    for p in sr:
        r2 = p.apply_exp_factor(r, complex(0.0, -0.01)) if p is sr[0] else p.apply_exp_factor(r2, complex(0.0, -0.01))
    assert_amps(out, vec(r2), msg="700-term sequence")
    eg, er = gpu.SumOp(sg).expectation_value(out), ref.SumOp(sr).expectation_value(r2)
    assert abs(eg - er) <= EXP_RTOL * max(1.0, abs(er)), (eg, er)
    # argument errors leave the state untouched and report the reference's variant
    bad = sg[:3] + [gpu.PauliString.new(1.0).with_op(n + 2, gpu.Pauli.X)]
    before = vec(out).copy()
    with pytest.raises(gpu.Error) as e:
        apply_exp_sequence_(out, bad, [1.0] * 4)
    assert e.value.variant == "InvalidQubitIndex"
    assert np.array_equal(vec(out), before)
    assert apply_exp_sequence_(out, [], []) is out


@pytest.mark.parametrize("absorb", [1, 0])
def test_cnot_absorption_matches_oracle(gpu, ref, absorb):
    """Option "absorb": a CNOT next to a single-qubit gate on its target is folded into two half-populated ops
    (G X / X G where the control is set, G where it is clear).  Random H / RX / RY / RZ / CNOT lists exercise both
    composition orders, register, tile and lane controls, and targets on lanes and registers."""
    gpu.engine.set_option("absorb", absorb)
    try:
        rng = np.random.default_rng(11)
        for trial in range(6):
            n = 9 + trial
            bg, br = gpu.CircuitBuilder(n), ref.CircuitBuilder(n)
            for _ in range(160):
                k = int(rng.integers(0, 6))
                t, c = [int(x) for x in rng.permutation(n)[:2]]
                a = float(rng.uniform(-3, 3))
                for b in (bg, br):
                    if k == 0: b.h_gate(t)
                    elif k == 1: b.rx_gate(t, a)
                    elif k == 2: b.ry_gate(t, a)
                    elif k == 3: b.cnot_gate(t, c)
                    elif k == 4: b.rz_gate(t, a)
                    else: b.y_gate(t)
            g, r = _pair(gpu, ref, n, seed=600 + trial)
            assert_amps(bg.build().execute(g), vec(br.build().execute(r)), msg=f"absorb={absorb} n={n}")
    finally:
        gpu.engine.set_option("absorb", 1)


def test_functional_api_clones_through_the_buffer_cache(gpu, ref):
    """The reference's API is functional (`&self -> State`): every gate call clones.  Clones come from a size-keyed
    device-buffer cache; states that die and are reborn in a tight loop must never alias live data."""
    n = 12
    g, r = _pair(gpu, ref, n, seed=700)
    keep = [g]
    for i in range(40):
        q = i % n
        g2 = g.h(q).rx((q + 3) % n, 0.1 * (i + 1)).cnot(q, (q + 1) % n)       # three clones, two die at once
        r = r.h(q).rx((q + 3) % n, 0.1 * (i + 1)).cnot(q, (q + 1) % n)
        if i % 7 == 0:
            keep.append(g2)                                                    # some stay alive across iterations
        g = g2
    assert_amps(g, vec(r), msg="chain of functional gate calls")
    first = vec(keep[0])
    assert_amps(keep[0], vec(ref.random_state(n, 700)), msg="the original state is untouched")
    gpu.engine.set_option("pool_mb", 0)            # trim and disable: same results without the cache
    try:
        h = keep[1].h(0).h(0)
        assert_amps(h, vec(keep[1]), tol=1e-14, msg="cache off")
    finally:
        gpu.engine.set_option("pool_mb", 4096)
    assert np.array_equal(vec(keep[0]), first)


@pytest.mark.parametrize("late", [0, 1])
def test_late_table_placement_both_settings_match_oracle(gpu, ref, late):
    """Option "late_tables" only moves unconditional phase-table ops inside a pass (host side, CPU-checked in
    tests/test_window_lowering.py); both settings must give the oracle's state on the device too."""
    from quant_iron_b200 import workloads as w
    n = 14
    specs = w.random_layered_circuit(n, 16)
    gpu.engine.set_option("late_tables", late)
    try:
        out = w.build_circuit(gpu, n, specs).execute(gpu.State.new_zero(n))
    finally:
        gpu.engine.set_option("late_tables", 1)
    want = w.build_circuit(ref, n, specs).execute(ref.State.new_zero(n))
    assert_amps(out, vec(want), msg=f"late_tables={late}")
