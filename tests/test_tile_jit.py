"""Circuit-specialised tile modules (csrc/tile_jit.cuh).

CPU tier: the PTX the generator writes for real circuits is assembled with the toolkit's `ptxas` for sm_100a (syntax,
register budget, no unresolved labels) -- the driver's assembler on the GPU box is the same code.
GPU tier: a circuit run on the assembled modules (option jit = 2) gives BIT-IDENTICAL amplitudes to the interpreting
kernel k_tile (jit = 0) -- the generator follows run_ops_tile instruction for instruction -- and both match the oracle
within the north-star tolerance (1e-12 max abs amplitude error).
"""
import concurrent.futures
import os
import shutil
import subprocess
import tempfile

import numpy as np
import pytest

import window_interp as wi
from conftest import assert_amps, vec


def _ptx_of(circuit, n):
    import quant_iron_b200 as gpu
    gpu.engine.set_option("debug_ptx", 1)
    try:
        blob = wi.lower(circuit, n)
    finally:
        gpu.engine.set_option("debug_ptx", 0)
    text = bytes(blob).decode("ascii", errors="replace")
    parts = text.split("//---PASS")[1:]
    out = []
    for p in parts:
        head, body = p.split("---\n", 1)
        end = body.rfind("}\n")
        out.append((int(head.split("coef=")[1]), body[:end + 2]))      # (coefficients, text)
    return out


def _assemble(texts):
    ptxas = shutil.which("ptxas") or "/usr/local/cuda/bin/ptxas"
    if not os.path.exists(ptxas):
        pytest.skip("no ptxas in this image")
    tmp = tempfile.mkdtemp(prefix="qi_ptx_")

    def one(i):
        src = os.path.join(tmp, f"p{i}.ptx")
        with open(src, "w") as f:
            f.write(texts[i])
        r = subprocess.run([ptxas, "-arch", "sm_100a", "-v", "-o", os.path.join(tmp, f"p{i}.cubin"), src], capture_output=True, text=True)
        return i, r.returncode, r.stderr
    try:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 4)) as ex:
            results = list(ex.map(one, range(len(texts))))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    for i, rc, err in results:
        assert rc == 0, f"pass {i}: ptxas failed\n{err[-2000:]}"
        assert "Used" in err and "registers" in err
        regs = int(err.split("Used ")[1].split(" registers")[0])
        assert regs <= 128, f"pass {i}: {regs} registers"
    return results


@pytest.fixture
def tile11():
    """k_tile / the modules take every state they can hold; tiles do not slide, so that the interpreter (jit = 0) runs the very
    schedule the modules are generated from (states that run on modules never slide: run_circuit_windowed)."""
    import quant_iron_b200 as gpu
    gpu.engine.set_option("tile_min_qubits", 11)
    gpu.engine.set_option("tile_slide", 0)
    yield gpu
    gpu.engine.set_option("tile_min_qubits", 18)
    gpu.engine.set_option("tile_slide", 1)


def test_benchmark_circuit_modules_assemble():
    """30 qubits, depth 40 (the bench workload): every tile pass becomes one module; the coefficient block fits the parameter space."""
    import quant_iron_b200 as gpu
    from quant_iron_b200 import workloads as w
    c = w.build_circuit(gpu, 30, w.random_layered_circuit(30, 40))
    mods = _ptx_of(c, 30)
    assert 10 <= len(mods) <= 30
    assert all(ncoef * 8 + 64 <= 32000 for ncoef, _ in mods)
    # the text is a function of the structure only: the same gates with slightly different angles reuse (nearly) every module
    # (a rotation whose cosine changes sign, or a pivot crossing the lifting threshold, changes the structure of its pass)
    specs2 = [(k, t, c_, [p + 1e-4 for p in ps]) for (k, t, c_, ps) in w.random_layered_circuit(30, 40)]
    mods2 = _ptx_of(w.build_circuit(gpu, 30, specs2), 30)
    assert len(mods2) == len(mods)
    assert sum(a[1] == b[1] for a, b in zip(mods, mods2)) >= len(mods) - 2
    assert any(a[1] == b[1] for a, b in zip(mods, mods2))
    _assemble([t for _, t in mods])


def test_qft_and_fuzz_modules_assemble(tile11):
    gpu = tile11
    from quant_iron_b200 import workloads as w
    texts = []
    for n in (12, 17, 30):
        texts += [t for _, t in _ptx_of(w.build_circuit(gpu, n, w.qft_specs(n)), n)]
    rng = np.random.default_rng(7)
    n = 13
    b = gpu.CircuitBuilder(n)
    for _ in range(200):
        kind = int(rng.integers(0, 12))
        t, c1, c2, t2 = [int(q) for q in rng.permutation(n)[:4]]
        ang = float(rng.uniform(-3, 3))
        nc = int(rng.integers(0, 3))
        ctrls = [c1, c2][:nc]
        if kind == 0: b.ch_gates([t], ctrls) if nc else b.h_gate(t)
        elif kind == 1: b.cx_gates([t], ctrls) if nc else b.x_gate(t)
        elif kind == 2: b.cy_gates([t], ctrls) if nc else b.y_gate(t)
        elif kind == 3: b.cz_gates([t], ctrls) if nc else b.z_gate(t)
        elif kind == 4: b.cs_gates([t], ctrls) if nc else b.t_gate(t)
        elif kind == 5: b.cp_gates([t], ctrls, ang) if nc else b.p_gate(t, ang)
        elif kind == 6: b.crx_gates([t], ctrls, ang) if nc else b.rx_gate(t, ang)
        elif kind == 7: b.cry_gates([t], ctrls, ang) if nc else b.ry_gate(t, ang)
        elif kind == 8: b.crz_gates([t], ctrls, ang) if nc else b.rz_gate(t, ang)
        elif kind == 9: b.toffoli_gate(c1, c2, t)
        elif kind == 10: b.ry_phase_gate(t, ang, 0.5 * ang)
        else: b.cnot_gate(t, c1)
    texts += [t for _, t in _ptx_of(b.build(), n)]
    assert len(texts) >= 8
    _assemble(texts)


# ---- GPU: assembled modules vs the interpreting kernel vs the oracle -----------------------------------------------------
def _run(gpu, circuit, state, jit):
    gpu.engine.set_option("jit", jit)
    try:
        gpu.engine.stats_reset()
        out = circuit.execute(state)
        sv = np.array(out.state_vector)
        return sv, gpu.engine.stats()
    finally:
        gpu.engine.set_option("jit", 1)


@pytest.mark.gpu
@pytest.mark.parametrize("n,depth", [(11, 8), (13, 12), (16, 20), (20, 16)])
def test_jit_layered_circuit_bit_identical_to_interpreter(gpu, ref, tile11, n, depth):
    from quant_iron_b200 import workloads as w
    specs = w.random_layered_circuit(n, depth, seed=300 + n)
    cg = w.build_circuit(gpu, n, specs)
    r0 = ref.random_state(n, 40 + n)
    sv_i, st_i = _run(gpu, cg, gpu.State(r0.state_vector, n), 0)
    sv_j, st_j = _run(gpu, cg, gpu.State(r0.state_vector, n), 2)
    assert st_i.get("gate_tile", {}).get("launches", 0) > 0 and "gate_tile_jit" not in st_i
    assert st_j.get("gate_tile_jit", {}).get("launches", 0) > 0 and st_j.get("gate_tile", {}).get("launches", 0) == 0, st_j
    assert gpu.engine.jit_stats()["failed"] == 0
    assert np.array_equal(sv_i, sv_j), f"max diff {np.max(np.abs(sv_i - sv_j)):.3e}"
    if n <= 16:
        out_r = w.build_circuit(ref, n, specs).execute(r0)
        assert np.max(np.abs(sv_j - vec(out_r))) <= 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("n,seed", [(11, 1), (12, 2), (13, 3), (14, 4)])
def test_jit_fuzzed_gate_lists_match_interpreter_and_oracle(gpu, ref, tile11, n, seed):
    """every operator kind under random controls (thread-, tile- and register-bit predicates, conditional swaps, merged tables)"""
    rng = np.random.default_rng(seed)
    bg, br = gpu.CircuitBuilder(n), ref.CircuitBuilder(n)
    for _ in range(220):
        kind = int(rng.integers(0, 14))
        t, c1, c2, t2 = [int(q) for q in rng.permutation(n)[:4]]
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
    r0 = ref.random_state(n, 500 + seed)
    cg = bg.build()
    sv_i, _ = _run(gpu, cg, gpu.State(r0.state_vector, n), 0)
    sv_j, st_j = _run(gpu, cg, gpu.State(r0.state_vector, n), 2)
    assert st_j.get("gate_tile_jit", {}).get("launches", 0) > 0
    assert gpu.engine.jit_stats()["failed"] == 0
    out_r = br.build().execute(r0)
    assert np.max(np.abs(sv_j - vec(out_r))) <= 1e-12
    assert np.max(np.abs(sv_i - sv_j)) <= 1e-15           # tables: the division may differ in the last bit; everything else is exact


@pytest.mark.gpu
@pytest.mark.parametrize("n", [12, 16, 20])
def test_jit_qft_closed_form_and_interpreter(gpu, ref, tile11, n):
    from quant_iron_b200 import workloads as w
    cg = w.build_circuit(gpu, n, w.qft_specs(n))
    sv_i, _ = _run(gpu, cg, gpu.State.new_plus(n), 0)
    sv_j, st_j = _run(gpu, cg, gpu.State.new_plus(n), 2)
    assert st_j.get("gate_tile_jit", {}).get("launches", 0) > 0
    assert abs(sv_j[0] - 1.0) <= 1e-12 and np.max(np.abs(sv_j[1:])) <= 1e-12
    assert np.max(np.abs(sv_i - sv_j)) <= 1e-15
    r0 = ref.random_state(n, 77) if n <= 16 else None
    if r0 is not None:
        a, _ = _run(gpu, cg, gpu.State(r0.state_vector, n), 2)
        out_r = w.build_circuit(ref, n, w.qft_specs(n)).execute(r0)
        assert np.max(np.abs(a - vec(out_r))) <= 1e-12


@pytest.mark.gpu
def test_jit_background_policy_switches_to_modules(gpu, ref, tile11):
    """jit = 1: the first execution of a structure runs on k_tile while its modules are assembled; after a drain the same
    circuit -- and the same structure with other angles -- runs on the modules, with identical amplitudes."""
    from quant_iron_b200 import workloads as w
    n = 14
    gpu.engine.set_option("jit_min_qubits", 11)
    try:
        specs = w.random_layered_circuit(n, 10, seed=909)
        cg = w.build_circuit(gpu, n, specs)
        r0 = ref.random_state(n, 5)
        before = gpu.engine.jit_stats()["modules"]
        sv1, st1 = _run(gpu, cg, gpu.State(r0.state_vector, n), 1)
        gpu.engine.jit_drain()
        assert gpu.engine.jit_stats()["modules"] > before and gpu.engine.jit_stats()["pending"] == 0
        sv2, st2 = _run(gpu, cg, gpu.State(r0.state_vector, n), 1)
        assert st2.get("gate_tile_jit", {}).get("launches", 0) > 0 and st2.get("gate_tile", {}).get("launches", 0) == 0
        assert np.array_equal(sv1, sv2)
        mods = gpu.engine.jit_stats()["modules"]
        _run(gpu, cg, gpu.State(r0.state_vector, n), 1)
        assert gpu.engine.jit_stats()["modules"] == mods          # cached
    finally:
        gpu.engine.set_option("jit_min_qubits", 24)


@pytest.mark.gpu
@pytest.mark.parametrize("option,value", [("jit_groups", 2), ("jit_groups", 4), ("jit_ctas", 3), ("jit_prefetch", 1), ("jit_prefetch", 2), ("jit_stage", 1)])
def test_jit_module_variants_bit_identical(gpu, ref, tile11, option, value):
    """The variants of the module skeleton -- CTAs that work on 2 or 4 tiles side by side, 3 CTAs per SM, L2 prefetch of the
    next tile, and the next tile staged into shared memory by bulk async copies (cp.async.bulk + mbarrier) -- move the data
    differently and compute exactly the same."""
    from quant_iron_b200 import workloads as w
    n = 17
    specs = w.random_layered_circuit(n, 14, seed=4242) + w.qft_specs(n)
    cg = w.build_circuit(gpu, n, specs)
    r0 = ref.random_state(n, 11)
    sv_1, _ = _run(gpu, cg, gpu.State(r0.state_vector, n), 2)
    gpu.engine.set_option(option, value)
    try:
        sv_g, st = _run(gpu, cg, gpu.State(r0.state_vector, n), 2)
        # a second, longer run on the same modules: many tiles per CTA (the staged variant's mbarrier phase flips every tile)
        n2 = 21
        c2 = w.build_circuit(gpu, n2, w.random_layered_circuit(n2, 6, seed=99))
        b_1 = None
        gpu.engine.set_option(option, {"jit_groups": 1, "jit_ctas": 4}.get(option, 0))
        b_1, _ = _run(gpu, c2, gpu.State.new_zero(n2), 2)
        gpu.engine.set_option(option, value)
        b_g, _ = _run(gpu, c2, gpu.State.new_zero(n2), 2)
    finally:
        gpu.engine.set_option(option, {"jit_groups": 1, "jit_ctas": 4}.get(option, 0))
    assert st.get("gate_tile_jit", {}).get("launches", 0) > 0
    assert gpu.engine.jit_stats()["failed"] == 0
    assert np.array_equal(sv_1, sv_g)
    assert np.array_equal(b_1, b_g)


@pytest.mark.gpu
@pytest.mark.parametrize("n,depth", [(14, 16), (16, 24), (21, 12)])
def test_jit_sliding_tiles_with_restored_layout(gpu, ref, n, depth):
    """option tile_restore: the modules run the SLIDING schedule (fewer passes) and relabel-only passes put every qubit back, so
    that the state keeps its layout and a second execution reuses every module."""
    from quant_iron_b200 import sharded, workloads as w
    specs = w.random_layered_circuit(n, depth, seed=77 + n)
    cg = w.build_circuit(gpu, n, specs)
    r0 = ref.random_state(n, 3)
    gpu.engine.set_option("tile_min_qubits", 11)
    gpu.engine.set_option("tile_restore", 1)
    gpu.engine.set_option("jit", 2)
    try:
        st = gpu.State(r0.state_vector, n)
        cg.execute_(st)
        phys, _ = sharded.layout(st)
        assert list(phys[:n]) == list(range(n)), phys[:n]
        mods = gpu.engine.jit_stats()["modules"]
        gpu.engine.stats_reset()
        cg.execute_(st)                                       # same layout -> same structures -> no new module
        assert gpu.engine.jit_stats()["modules"] == mods
        assert gpu.engine.stats().get("gate_tile_jit", {}).get("launches", 0) > 0
        st2 = gpu.State(r0.state_vector, n)
        cg.execute_(st2)
        sv = np.array(st2.state_vector)
    finally:
        gpu.engine.set_option("jit", 1)
        gpu.engine.set_option("tile_restore", 0)
        gpu.engine.set_option("tile_min_qubits", 18)
    assert gpu.engine.jit_stats()["failed"] == 0
    if n <= 16:
        out_r = w.build_circuit(ref, n, specs).execute(r0)
        assert np.max(np.abs(sv - vec(out_r))) <= 1e-12
    else:
        gpu.engine.set_option("jit", 0)
        try:
            st3 = gpu.State(r0.state_vector, n)
            cg.execute_(st3)
            assert np.max(np.abs(sv - np.array(st3.state_vector))) <= 1e-14
        finally:
            gpu.engine.set_option("jit", 1)
