"""The host half of the fused window executor, checked WITHOUT a GPU: `qi_debug_lower` serialises the device programs
`qi_apply_circuit` would launch (scheduling, CNOT absorption, merged phase tables, slot masks, tile predicates, lazy SWAP
relabelling, per-chunk / per-shard gate translation) and tests/window_interp.py interprets them with numpy.  The result
must match the oracle to the north-star bar.  (The CUDA half -- that the kernels do what the ops mean -- is the `-m gpu`
parity suite.)"""
import numpy as np
import pytest

import window_interp as wi
from conftest import AMP_TOL, vec
from test_host_pipeline import _fuzz_builders, _plan


def _run(circuit, n, start, regs=0):
    v = np.array(start, dtype=np.complex128)
    phys, steps = wi.execute(wi.lower(circuit, n, regs=regs), v, n)
    return wi.to_logical(v, n, phys), steps


@pytest.mark.parametrize("n,seed,regs", [(8, 1, 3), (9, 2, 4), (10, 3, 4), (10, 4, 5), (11, 5, 4), (12, 6, 4), (11, 7, 5), (9, 8, 3)])
def test_fuzzed_gate_lists_lowered_programs_match_oracle(ref, n, seed, regs):
    import quant_iron_b200 as gpu
    cg, cr = _fuzz_builders([gpu, ref], n, seed, count=200, lazy_swaps=(seed % 2 == 0))
    start = ref.random_state(n, 40 + seed)
    got, steps = _run(cg, n, start.state_vector, regs)
    want = vec(cr.execute(start))
    assert float(np.max(np.abs(got - want))) <= AMP_TOL
    assert any(s[0] == "pass" for s in steps)


@pytest.mark.parametrize("cz", [0, 1])
@pytest.mark.parametrize("n,depth", [(10, 12), (13, 10), (14, 6)])
def test_layered_circuit_lowered_programs_match_oracle(ref, n, depth, cz):
    import quant_iron_b200 as gpu
    from quant_iron_b200 import workloads as w
    specs = w.random_layered_circuit(n, depth)
    start = ref.random_state(n, 3)
    gpu.engine.set_option("cz_rewrite", cz)
    try:
        got, steps = _run(w.build_circuit(gpu, n, specs), n, start.state_vector)
    finally:
        gpu.engine.set_option("cz_rewrite", 1)
    want = vec(w.build_circuit(ref, n, specs).execute(start))
    assert float(np.max(np.abs(got - want))) <= AMP_TOL
    ops = np.concatenate([s[3] for s in steps if s[0] == "pass"])
    assert (ops["kind"] == wi.WK_TABLE).any()                      # merged RZ runs
    if not cz:
        assert ((ops["kind"] <= wi.LAST_PAIR) & (ops["c_tval"] != ops["c_tile"])).any() or n < 12    # negative controls of absorbed CNOTs


def test_cx_next_to_h_becomes_cz_bit_exact(ref):
    """[CX(c,t), H(t)] == [H(t), CZ(c,t)] and [H(t), CX(c,t)] == [CZ(c,t), H(t)] hold bit for bit in the reference's arithmetic;
    the executor uses them (option cz_rewrite) to turn register-swap ops into phase-table members.  The rewritten programs
    must reproduce the programs without the rewrite EXACTLY, and hold fewer X ops."""
    import quant_iron_b200 as gpu
    from quant_iron_b200 import workloads as w
    n = 13
    specs = w.random_layered_circuit(n, 14)
    start = ref.random_state(n, 21)
    out, nx = {}, {}
    for cz in (0, 1):
        gpu.engine.set_option("cz_rewrite", cz)
        gpu.engine.set_option("absorb", 0)
        try:
            out[cz], steps = _run(w.build_circuit(gpu, n, specs), n, start.state_vector)
        finally:
            gpu.engine.set_option("cz_rewrite", 1)
            gpu.engine.set_option("absorb", 1)
        ops = np.concatenate([s[3] for s in steps if s[0] == "pass"])
        nx[cz] = int((ops["kind"] == wi.WK_X).sum())
    assert nx[1] < 0.6 * nx[0], nx
    assert float(np.max(np.abs(out[1] - out[0]))) <= 4e-16          # same products and sums; merged table phases differ by an ulp
    want = vec(w.build_circuit(ref, n, specs).execute(start))
    assert float(np.max(np.abs(out[1] - want))) <= AMP_TOL


@pytest.mark.parametrize("n", [9, 12])
def test_qft_lowered_programs_match_closed_form(ref, n):
    """QFT|+..+> = |0..0> through the phase-table ops (one per Hadamard) and the relabelled final swaps."""
    import quant_iron_b200 as gpu
    from quant_iron_b200 import workloads as w
    c = w.build_circuit(gpu, n, w.qft_specs(n))
    plus = np.full(1 << n, 1.0 / np.sqrt(float(1 << n)), dtype=np.complex128)
    got, steps = _run(c, n, plus)
    assert abs(got[0] - 1.0) <= AMP_TOL and float(np.max(np.abs(got[1:]))) <= AMP_TOL
    start = ref.random_state(n, 9)
    got, _ = _run(c, n, start.state_vector)
    want = vec(w.build_circuit(ref, n, w.qft_specs(n)).execute(start))
    assert float(np.max(np.abs(got - want))) <= AMP_TOL


@pytest.fixture
def lean():
    import quant_iron_b200 as gpu
    gpu.engine.set_option("lean", 1)
    yield
    gpu.engine.set_option("lean", 0)


@pytest.mark.parametrize("n,seed,regs", [(9, 11, 3), (10, 12, 4), (11, 13, 4), (12, 14, 5), (10, 15, 4)])
def test_lean_lowering_fuzz_matches_oracle(ref, lean, n, seed, regs):
    """Option "lean": unit-form H / RX / real 2x2 ops with ONE deferred scale per pass (folded into a phase table or a
    scale op) -- every operator kind, controls, absorbed CNOTs, relabelled SWAPs."""
    import quant_iron_b200 as gpu
    cg, cr = _fuzz_builders([gpu, ref], n, seed, count=240, lazy_swaps=(seed % 2 == 0))
    start = ref.random_state(n, 40 + seed)
    got, steps = _run(cg, n, start.state_vector, regs)
    assert float(np.max(np.abs(got - vec(cr.execute(start))))) <= AMP_TOL


@pytest.mark.parametrize("n,depth", [(12, 12), (14, 8)])
def test_lean_lowering_layered_circuit(ref, lean, n, depth):
    """The benchmark generator: most H / RX gates (and absorbed-CNOT pairs) take the lean forms; the scale rides on the
    pass's RZ phase table where there is one."""
    import quant_iron_b200 as gpu
    from quant_iron_b200 import workloads as w
    specs = w.random_layered_circuit(n, depth)
    start = ref.random_state(n, 3)
    got, steps = _run(w.build_circuit(gpu, n, specs), n, start.state_vector)
    want = vec(w.build_circuit(ref, n, specs).execute(start))
    assert float(np.max(np.abs(got - want))) <= AMP_TOL
    ops = np.concatenate([s[3] for s in steps if s[0] == "pass"])
    n_lean = int(((ops["kind"] >= wi.WK_REALUP) & (ops["kind"] <= wi.WK_RXSU)).sum())
    n_scaled = int(((ops["kind"] >= wi.WK_RX) & (ops["kind"] <= wi.WK_REAL)).sum())
    assert n_lean > 2 * n_scaled, (n_lean, n_scaled)
    assert int((ops["kind"] == wi.WK_SCALE).sum()) <= len([s for s in steps if s[0] == "pass"])


@pytest.mark.parametrize("absorb,fuse", [(0, 1), (1, 0), (0, 0)])
def test_lowering_options(ref, absorb, fuse):
    import quant_iron_b200 as gpu
    from quant_iron_b200 import workloads as w
    n = 11
    specs = w.random_layered_circuit(n, 8)
    start = ref.random_state(n, 4)
    want = vec(w.build_circuit(ref, n, specs).execute(start))
    gpu.engine.set_option("absorb", absorb)
    gpu.engine.set_option("fuse", fuse)
    try:
        got, steps = _run(w.build_circuit(gpu, n, specs), n, start.state_vector)
    finally:
        gpu.engine.set_option("absorb", 1)
        gpu.engine.set_option("fuse", 1)
    assert float(np.max(np.abs(got - want))) <= AMP_TOL
    if not fuse:
        assert len(steps) == len(specs)


@pytest.mark.parametrize("n,k,seed", [(11, 1, 1), (12, 2, 2), (12, 3, 3), (13, 3, 4)])
def test_host_pipeline_chunk_programs_match_oracle(ref, n, k, seed):
    """qi_execute_host on the CPU: front gates chunk by chunk (each chunk lowered as shard `c` of 2^k), middle on the
    whole vector, back gates chunk by chunk -- the exact programs the device would run -- against the oracle."""
    import quant_iron_b200 as gpu
    cg, cr = _fuzz_builders([gpu, ref], n, seed, count=160)
    order, nf, nm, nb = _plan(cg, n, k)
    start = ref.random_state(n, 70 + seed)
    want = vec(cr.execute(start))
    v = np.array(start.state_vector, dtype=np.complex128)
    clen = 1 << (n - k)

    def sub(idx):
        return gpu.Circuit.with_gates([cg.gates[i] for i in idx], n)

    def per_chunk(idx):
        if not idx:
            return
        c = sub(idx)
        for ch in range(1 << k):
            view = v[ch * clen:(ch + 1) * clen]
            phys, _ = wi.execute(wi.lower(c, n, rank=ch, world=1 << k), view, n - k)
            assert phys[:n] == list(range(n))

    per_chunk(order[:nf])
    if nm:
        phys, _ = wi.execute(wi.lower(sub(order[nf:nf + nm]), n), v, n)
        assert phys[:n] == list(range(n))
    per_chunk(order[nf + nm:])
    assert float(np.max(np.abs(v - want))) <= AMP_TOL
    assert nf > 0 and nb > 0


def _random_strings(apis, n, seed, count):
    rng = np.random.default_rng(seed)
    out = [[] for _ in apis]
    factors = []
    for trial in range(count):
        k = int(rng.integers(0, 8)) if trial != 7 else 0
        qs = [int(q) for q in rng.choice(n, size=min(k, n), replace=False)]
        c = complex(rng.uniform(-1, 1), rng.uniform(-0.3, 0.3) if trial % 3 == 0 else 0.0)
        which = [int(rng.integers(0, 3)) for _ in qs]
        for a, lst in zip(apis, out):
            p = a.PauliString.new(c)
            for q, w_ in zip(qs, which):
                p.add_op(q, [a.Pauli.X, a.Pauli.Y, a.Pauli.Z][w_])
            lst.append(p)
        factors.append(complex(rng.uniform(-0.2, 0.2), rng.uniform(-0.5, 0.5)) if trial % 4 == 0 else complex(0.0, rng.uniform(-0.5, 0.5)))
    return out, factors


@pytest.mark.parametrize("n,seed", [(9, 0), (10, 1), (11, 2), (12, 3)])
def test_pauli_exp_sequence_lowered_programs_match_oracle(ref, n, seed):
    """Random Pauli strings (X/Y/Z anywhere, real / imaginary / complex exponents, an empty string): the fused
    Pauli-window passes and the terms that run alone, interpreted on the CPU, vs one apply_exp_factor per string."""
    import quant_iron_b200 as gpu
    (sg, sr), factors = _random_strings([gpu, ref], n, seed, 80)
    r = ref.random_state(n, 300 + seed)
    v = np.array(r.state_vector, dtype=np.complex128)
    for p, f in zip(sr, factors):
        r = p.apply_exp_factor(r, f)
    passes, singles = wi.execute_pauli(wi.lower_pauli(sg, factors, n), v, n)
    nrm = max(1.0, float(np.sqrt(np.vdot(vec(r), vec(r)).real)))
    assert float(np.max(np.abs(v - vec(r)))) <= AMP_TOL * nrm
    assert passes >= 1 and singles >= 1


def test_heisenberg_trotter_lowered_programs_match_oracle(ref):
    """Two first-order Trotter steps of the 12-site Heisenberg chain (heisenberg.rs:68-96 term order) as one fused
    sequence: every term is fusable, terms share passes, and the interpreted programs track the oracle."""
    import quant_iron_b200 as gpu
    n, dt = 12, 0.01
    hg, hr = gpu.heisenberg_1d(n, 1.0, 2.0, 3.0, 0.5, 0.1), ref.heisenberg_1d(n, 1.0, 2.0, 3.0, 0.5, 0.1)
    r = ref.State.new_plus(n)
    v = np.array(r.state_vector, dtype=np.complex128)
    r = ref.trotter_evolve_state(hr, r, dt, 2, ref.TrotterOrder.First)
    strings = list(hg.terms) * 2
    passes, singles = wi.execute_pauli(wi.lower_pauli(strings, [complex(0.0, -dt)] * len(strings), n), v, n)
    assert singles == 0 and passes < len(strings) // 4
    assert float(np.max(np.abs(v - vec(r)))) <= AMP_TOL


def test_late_table_placement_fewer_diagonal_ops_same_state(ref):
    """Unconditional phase tables are placed as late as their most urgent member allows (minimum piercing set of the
    gates' commutation intervals): the layered circuit needs one diagonal op in almost every pass instead of up to six,
    and the state is the same."""
    import quant_iron_b200 as gpu
    from quant_iron_b200 import workloads as w
    n = 13
    specs = w.random_layered_circuit(n, 16)
    start = ref.random_state(n, 8)
    want = vec(w.build_circuit(ref, n, specs).execute(start))
    counts = {}
    gpu.engine.set_option("cz_rewrite", 0)          # counts below: RZ tables only, no controlled-Z members
    for late in (0, 1):
        gpu.engine.set_option("late_tables", late)
        try:
            got, steps = _run(w.build_circuit(gpu, n, specs), n, start.state_vector)
        finally:
            gpu.engine.set_option("late_tables", 1)
        assert float(np.max(np.abs(got - want))) <= AMP_TOL
        ops = np.concatenate([s[3] for s in steps if s[0] == "pass"])
        counts[late] = (int(((ops["kind"] == wi.WK_TABLE) | (ops["kind"] == wi.WK_RZ)).sum()), len(steps))
    assert counts[1][1] == counts[0][1]                       # same passes
    assert counts[1][0] < 0.8 * counts[0][0], counts          # fewer diagonal ops
    # the benchmark circuit on the warp-tile executor: at most 3 diagonal ops in any pass (6 without), 123 instead of 197 in total
    gpu.engine.set_option("tile", 0)
    try:
        steps, _, _ = wi.parse(wi.lower(w.build_circuit(gpu, 30, w.random_layered_circuit(30, 40)), 30))
    finally:
        gpu.engine.set_option("tile", 1)
        gpu.engine.set_option("cz_rewrite", 1)
    per = [int(((s[3]["kind"] == wi.WK_TABLE) | (s[3]["kind"] == wi.WK_RZ)).sum()) for s in steps if s[0] == "pass"]
    assert max(per) <= 3 and sum(per) <= 130, (max(per), sum(per))


@pytest.mark.parametrize("n,seed", [(9, 0), (11, 1), (12, 2)])
def test_expectation_groups_match_oracle(ref, n, seed):
    """SumOp::expectation_value batching: first-fit groups of terms sharing one register window, interpreted on the CPU,
    plus the terms left to the per-term kernel (evaluated here with the oracle), vs the oracle's term-by-term sum."""
    import quant_iron_b200 as gpu
    (sg, sr), _ = _random_strings([gpu, ref], n, 50 + seed, 70)
    psi = ref.random_state(n, 60 + seed)
    v = np.array(psi.state_vector, dtype=np.complex128)
    got, ngroups, left = wi.expect_groups(wi.lower_expect(sg, n), v, n)
    for i in left:
        got += ref.SumOp([sr[i]]).expectation_value(psi)
    want = ref.SumOp(sr).expectation_value(psi)
    assert abs(got - want) <= 1e-10 * max(1.0, abs(want))
    assert 1 <= ngroups < len(sg) - len(left)


def test_heisenberg_expectation_groups(ref):
    """The 96 terms of the 24-site chain (BASELINE config 3) share a handful of read passes; value checked at 12 sites."""
    import quant_iron_b200 as gpu
    n = 12
    hg, hr = gpu.heisenberg_1d(n, 1.0, 2.0, 3.0, 0.5, 0.1), ref.heisenberg_1d(n, 1.0, 2.0, 3.0, 0.5, 0.1)
    psi = ref.random_state(n, 77)
    got, ngroups, left = wi.expect_groups(wi.lower_expect(list(hg.terms), n), np.array(psi.state_vector), n)
    want = hr.expectation_value(psi)
    assert not left and ngroups <= 6
    assert abs(got - want) <= 1e-10 * max(1.0, abs(want))


# ---- CTA-tile passes (k_tile): rounds of 4 register qubits out of 11 tile qubits ---------------------------------------
@pytest.fixture
def tile11():
    """Force the CTA-tile executor for every state with >= 11 local qubits (default: >= 18)."""
    import quant_iron_b200 as gpu
    gpu.engine.set_option("tile_min_qubits", 11)
    yield
    gpu.engine.set_option("tile_min_qubits", 18)


def _tile_steps(steps):
    return [s for s in steps if s[0] == "tile"]


@pytest.mark.parametrize("n,seed", [(11, 1), (12, 2), (12, 3), (13, 4), (14, 5), (13, 6), (11, 7), (12, 8)])
def test_tile_fuzzed_gate_lists_match_oracle(ref, tile11, n, seed):
    import quant_iron_b200 as gpu
    cg, cr = _fuzz_builders([gpu, ref], n, seed, count=260, lazy_swaps=(seed % 2 == 0))
    start = ref.random_state(n, 140 + seed)
    got, steps = _run(cg, n, start.state_vector)
    want = vec(cr.execute(start))
    assert float(np.max(np.abs(got - want))) <= AMP_TOL
    assert _tile_steps(steps) and not any(s[0] == "pass" for s in steps)


@pytest.mark.parametrize("n,depth,lean", [(12, 12, 0), (14, 10, 0), (15, 8, 0), (13, 10, 1), (15, 6, 1)])
def test_tile_layered_circuit_matches_oracle(ref, tile11, n, depth, lean):
    import quant_iron_b200 as gpu
    from quant_iron_b200 import workloads as w
    specs = w.random_layered_circuit(n, depth)
    start = ref.random_state(n, 5)
    gpu.engine.set_option("lean", lean)
    try:
        got, steps = _run(w.build_circuit(gpu, n, specs), n, start.state_vector)
    finally:
        gpu.engine.set_option("lean", 0)
    want = vec(w.build_circuit(ref, n, specs).execute(start))
    assert float(np.max(np.abs(got - want))) <= AMP_TOL
    tiles = _tile_steps(steps)
    assert tiles
    # the point of the tile pass: many rounds share one HBM pass, and gates on qubits 0..4 are register gates like any other
    assert max(len(t[2]) for t in tiles) >= 3
    kinds = np.concatenate([ops["kind"] for t in tiles for (_, _, ops) in t[2] if len(ops)])
    assert (kinds == wi.WK_TABLE).any()
    if lean:
        assert (kinds >= wi.WK_REALUP).any()


@pytest.mark.parametrize("n", [11, 14, 16])
def test_tile_qft_matches_closed_form_and_oracle(ref, tile11, n):
    import quant_iron_b200 as gpu
    from quant_iron_b200 import workloads as w
    c = w.build_circuit(gpu, n, w.qft_specs(n))
    plus = np.full(1 << n, 1.0 / np.sqrt(float(1 << n)), dtype=np.complex128)
    got, steps = _run(c, n, plus)
    assert abs(got[0] - 1.0) <= AMP_TOL and float(np.max(np.abs(got[1:]))) <= AMP_TOL
    assert _tile_steps(steps)
    start = ref.random_state(n, 19)
    got, _ = _run(c, n, start.state_vector)
    want = vec(w.build_circuit(ref, n, w.qft_specs(n)).execute(start))
    assert float(np.max(np.abs(got - want))) <= AMP_TOL


def test_tile_pass_count_of_the_benchmark_circuit():
    """30 qubits, depth 40: chosen, sliding 11-qubit tiles need a fraction of the 126 HBM passes of the warp-tile executor."""
    import quant_iron_b200 as gpu
    from quant_iron_b200 import workloads as w
    c = w.build_circuit(gpu, 30, w.random_layered_circuit(30, 40))
    gpu.engine.set_option("jit", 0)
    try:
        steps, _, _ = wi.parse(wi.lower(c, 30))
    finally:
        gpu.engine.set_option("jit", 1)
    tiles = _tile_steps(steps)
    assert len(steps) - len(tiles) <= 2
    assert len(tiles) <= 30, len(tiles)
    assert sum(len(ops) for t in tiles for (_, _, ops) in t[2]) >= 1300
    assert any(t[3] != t[1] for t in tiles)                # tiles slide: some pass stores its qubits in a new order
    # states that run on JIT modules (>= jit_min_qubits) keep their layout, so that the next execution of the circuit has the
    # same pass structures and reuses every module: more passes, none of them relabelling
    steps, _, _ = wi.parse(wi.lower(c, 30))
    tiles = _tile_steps(steps)
    assert len(tiles) <= 40, len(tiles)
    assert all(t[3] == t[1] for t in tiles)


@pytest.mark.parametrize("options", [{"tile_pform": 0}, {"tile_pform": 2}, {"tile_carry": 0}, {"tile_lean": 0}, {"tile_lean": 0, "tile_pform": 0, "tile_carry": 0}],
                         ids=lambda o: ",".join(f"{k}={v}" for k, v in o.items()))
def test_tile_lowering_options_match_oracle(ref, tile11, options):
    """Round-2 lowering options of the tile passes -- unit-form pair ops, P-form phase ops instead of small tables, the scalar
    carried across launches -- each against the oracle, with the remaining options at their defaults."""
    import quant_iron_b200 as gpu
    from quant_iron_b200 import workloads as w
    n = 14
    specs = w.random_layered_circuit(n, 14, seed=31) + w.qft_specs(n) + w.random_layered_circuit(n, 4, seed=32)
    start = ref.random_state(n, 8)
    defaults = {"tile_pform": 4, "tile_carry": 1, "tile_lean": 1}
    for k, v in options.items():
        gpu.engine.set_option(k, v)
    try:
        got, steps = _run(w.build_circuit(gpu, n, specs), n, start.state_vector)
    finally:
        for k, v in defaults.items():
            gpu.engine.set_option(k, v)
    want = vec(w.build_circuit(ref, n, specs).execute(start))
    assert float(np.max(np.abs(got - want))) <= AMP_TOL
    kinds = np.concatenate([ops["kind"] for t in _tile_steps(steps) for (_, _, ops) in t[2] if len(ops)])
    if options.get("tile_lean", 1) == 0:
        assert not np.isin(kinds, [wi.WK_REALUP, wi.WK_REALUM, wi.WK_RXU]).any()
    if options.get("tile_carry", 1) == 1 and options.get("tile_lean", 1) == 1:
        assert (kinds == wi.WK_SCALE).sum() <= 2              # one scalar per run (QFT swaps split the run in two at most)


@pytest.mark.parametrize("n,depth", [(17, 10), (19, 8)])
def test_tile_restore_relabel_passes_match_oracle(ref, n, depth):
    """option tile_restore (states that run on modules): sliding tiles, then relabel-only passes that put every qubit back --
    the final layout is the initial one and the amplitudes match the oracle."""
    import quant_iron_b200 as gpu
    from quant_iron_b200 import workloads as w
    specs = w.random_layered_circuit(n, depth, seed=5 + n)
    start = ref.random_state(n, 2)
    for k, v in {"tile_min_qubits": 11, "jit_min_qubits": 11, "tile_restore": 1}.items():
        gpu.engine.set_option(k, v)
    try:
        blob = wi.lower(w.build_circuit(gpu, n, specs), n)
    finally:
        for k, v in {"tile_min_qubits": 18, "jit_min_qubits": 24, "tile_restore": 0}.items():
            gpu.engine.set_option(k, v)
    steps, _, final_phys = wi.parse(blob)
    tiles = _tile_steps(steps)
    assert any(t[3] != t[1] for t in tiles)                                  # tiles slide ...
    assert list(final_phys[:n]) == list(range(n))                            # ... and the layout comes back
    v = np.array(start.state_vector)
    out_phys, _ = wi.execute(blob, v, n)
    got = wi.to_logical(v, n, out_phys)
    want = vec(w.build_circuit(ref, n, specs).execute(start))
    assert float(np.max(np.abs(got - want))) <= AMP_TOL
