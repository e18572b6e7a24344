"""Generate tests/golden/golden_r01.json from the CPU oracle.

The reference is a Rust crate that cannot be built in this image (no cargo/rustc), so these vectors are NOT
outputs of the reference binary: they are outputs of `oracle/` (the restatement pinned by the reference's own
known-answer tests, tests/test_ref_ported_*.py), frozen so that (a) any later change of the oracle's arithmetic
shows up as a diff against committed numbers and (b) the GPU engine is compared with numbers that were fixed
before it was tuned.  Floats are stored as C99 hex strings (exact round trip).

    python tests/golden/make_golden.py            # rewrite the fixture
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_r01.json")


def hexvec(v):
    return [[float(z.real).hex(), float(z.imag).hex()] for z in v]


def cases(api, w):
    """Every case as (name, callable returning a dict of results) on an API with the reference's shape."""
    out = {}
    n = 10
    specs = w.random_layered_circuit(n, 8)
    st = w.build_circuit(api, n, specs).execute(api.State.new_zero(n))
    out["layered_n10_depth8"] = {"amps": hexvec(st.state_vector)}
    out["layered_n10_sample_q035_256shots_seed20260003"] = {"bins": [int(b) for b in st.sample_counts([0, 3, 5], 256, 20260003)]}
    mr = st.measure(api.MeasurementBasis.Computational, [1, 4, 7], seed=5)
    out["layered_n10_measure_q147_seed5"] = {"outcomes": [int(o) for o in mr.get_outcomes()], "amps": hexvec(mr.get_new_state().state_vector)}
    n = 8
    rs = api.random_state(n, 20260002) if hasattr(api, "random_state") else None
    if rs is None:
        rs = api.State.new_random(n, 20260002)
    qs = list(range(n))
    q = api.CircuitBuilder(n).add_subroutine(api.Subroutine.qft(qs, n)).build().execute(rs)
    out["qft_n8_random_state"] = {"amps": hexvec(q.state_vector)}
    h = api.heisenberg_1d(n, 1.0, 2.0, 3.0, 0.5, 0.1)
    for order, tag in ((api.TrotterOrder.First, "first"), (api.TrotterOrder.Second, "second")):
        ev = api.trotter_evolve_state(h, api.State.new_plus(n), 0.01, 3, order)
        e = h.expectation_value(ev)
        out[f"heisenberg_n8_trotter3_{tag}"] = {"amps": hexvec(ev.state_vector), "expectation": [float(e.real).hex(), float(e.imag).hex()]}
    ps = api.PauliString.new(complex(0.3, -0.2)).with_op(0, api.Pauli.X).with_op(3, api.Pauli.Y).with_op(6, api.Pauli.Z)
    out["pauli_x0y3z6_exp_factor"] = {"amps": hexvec(ps.apply_exp_factor(rs, complex(0.1, -0.25)).state_vector)}
    return out


def main():
    from oracle import refapi as ref
    from quant_iron_b200 import workloads as w
    data = {"generator": "tests/golden/make_golden.py", "source": "oracle/refapi.py + oracle/qi_oracle.c (round 1)",
            "uniform_seed20260003_first8": [float(ref.uniform(20260003, k)).hex() for k in range(8)],
            "cases": cases(ref, w)}
    with open(OUT, "w") as f:
        json.dump(data, f, indent=0)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
