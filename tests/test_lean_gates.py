"""Option "lean" of the window executor (unit-form H / RX / real 2x2 ops with one deferred scale per pass; off by
default, built after the round's GPU budget was spent): the CUDA half against the oracle.  The host half (which ops are
converted, the scale folding) is checked on the CPU tier in tests/test_window_lowering.py."""
import numpy as np
import pytest

from conftest import AMP_TOL, assert_amps, vec
from test_host_pipeline import _fuzz_builders

pytestmark = [pytest.mark.gpu, pytest.mark.unproven]


@pytest.fixture
def lean(gpu):
    gpu.engine.set_option("lean", 1)
    yield
    gpu.engine.set_option("lean", 0)
    gpu.engine.set_option("prefetch", 0)
    gpu.engine.set_option("window_regs", 0)


@pytest.mark.parametrize("n,seed,regs", [(9, 21, 3), (10, 22, 4), (11, 23, 4), (12, 24, 5), (13, 25, 4), (12, 26, 4)])
def test_lean_fuzz_vs_oracle(gpu, ref, lean, n, seed, regs):
    gpu.engine.set_option("window_regs", regs)
    cg, cr = _fuzz_builders([gpu, ref], n, seed, count=240, lazy_swaps=(seed % 2 == 0))
    start = ref.random_state(n, 500 + seed)
    out = cg.execute(gpu.State(start.state_vector, n))
    assert_amps(out, vec(cr.execute(start)), msg=f"lean fuzz n={n} seed={seed}")


@pytest.mark.parametrize("n,depth", [(12, 20), (16, 40), (20, 12)])
def test_lean_layered_circuit_vs_oracle(gpu, ref, lean, n, depth):
    from quant_iron_b200 import workloads as w
    specs = w.random_layered_circuit(n, depth)
    out = w.build_circuit(gpu, n, specs).execute(gpu.State.new_zero(n))
    want = w.build_circuit(ref, n, specs).execute(ref.State.new_zero(n))
    assert_amps(out, vec(want), msg=f"lean layered n={n}")


def test_lean_matches_default_at_26_qubits(gpu, lean):
    from quant_iron_b200 import workloads as w
    n = 26
    c = w.build_circuit(gpu, n, w.random_layered_circuit(n, 10))
    a = c.execute(gpu.State.new_zero(n))
    gpu.engine.set_option("lean", 0)
    b = c.execute(gpu.State.new_zero(n))
    assert abs(a.norm_sqr() - 1.0) <= 1e-10
    ip = a.inner_product(b)
    assert abs(ip - 1.0) <= 1e-10
    for i in (0, 1, 12345, (1 << n) - 1):
        assert abs(a.amplitude(i) - b.amplitude(i)) <= AMP_TOL


@pytest.mark.parametrize("n,depth", [(14, 20), (22, 12)])
def test_lean_with_l2_prefetch_vs_default(gpu, ref, lean, n, depth):
    """Option "prefetch" (lean instantiations only) is a pure cache hint: the state must be bit-identical to lean alone."""
    from quant_iron_b200 import workloads as w
    c = w.build_circuit(gpu, n, w.random_layered_circuit(n, depth))
    a = vec(c.execute(gpu.State.new_zero(n)))
    gpu.engine.set_option("prefetch", 1)
    b = vec(c.execute(gpu.State.new_zero(n)))
    assert np.array_equal(a, b)
    if n <= 16:
        want = vec(w.build_circuit(ref, n, w.random_layered_circuit(n, depth)).execute(ref.State.new_zero(n)))
        assert float(np.max(np.abs(b - want))) <= AMP_TOL
