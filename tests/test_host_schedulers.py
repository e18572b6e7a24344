"""Host-only logic of the product (no GPU): the fused-pass scheduler, the Pauli-exp scheduler and the exchange
planners are pure functions of the gate / term list, exposed through debug entry points of the C ABI
(qi_debug_schedule, qi_debug_pauli_schedule, qi_shard_plan, qi_shard_plan_pauli)."""
import ctypes as C
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def _schedule(n, specs):
    from show_schedule import schedule
    return schedule(n, specs)


def test_qft_is_one_phase_table_per_hadamard():
    """The QFT's H + controlled-phase ladder merges into one phase-table op per Hadamard: a 33-qubit QFT
    (561 gates + 16 relabelled swaps) is 7 fused passes, a 20-qubit one 4."""
    from quant_iron_b200 import workloads as w
    for n, passes in ((33, 7), (20, 4)):
        rows = _schedule(n, w.qft_specs(n))
        assert len(rows) == passes, (n, len(rows))
        assert sum(r[0] for r in rows) == 0                      # no per-gate-kernel steps
        pair_ops = sum(r[2] + r[3] for r in rows)
        assert pair_ops == n                                     # exactly the n Hadamards remain as pair gates
        assert sum(r[5] for r in rows) >= n - 2                  # and (almost) every ladder became one table op


def test_layered_circuit_schedule_is_stable():
    """The benchmark circuit (30 q, depth 40, 1780 gates): 126 passes, every pass uses the 4 window qubits, most
    CNOTs are absorbed into their neighbours.  A change here moves the headline number: look before accepting it."""
    from quant_iron_b200 import workloads as w
    specs = w.random_layered_circuit(30, 40)
    assert len(specs) == 1780
    rows = _schedule(30, specs)
    assert 120 <= len(rows) <= 132, len(rows)
    assert all(r[1] == 4 for r in rows[:-1])
    assert sum(r[6] for r in rows) >= 380                        # absorbed CNOTs (580 in the circuit)
    heavy = [r for r in rows if r[2] >= 6]
    assert len(heavy) <= 30                                      # lane-qubit gates stay concentrated in few passes


def test_pauli_schedule_heisenberg_24():
    """BASELINE config 3: 96 terms x 50 first-order steps = 4800 exps in 300 fused passes (6 per step)."""
    import quant_iron_b200 as qi
    from quant_iron_b200 import _ffi
    h = qi.heisenberg_1d(24, 1.0, 2.0, 3.0, 0.5, 0.1)
    arr, n, keep = h.term_array()
    rows = (C.c_int32 * 4096)()
    cnt = C.c_uint64(0)
    _ffi.check(_ffi.lib.qi_debug_pauli_schedule(24, arr, n, 50, rows, 4096, C.byref(cnt)))
    per = [rows[i] for i in range(cnt.value)]
    assert sum(per) == 4800 and 0 not in per
    assert cnt.value <= 320, cnt.value


@pytest.mark.parametrize("n,world", [(31, 2), (32, 4), (33, 8), (28, 8)])
def test_gate_planner_layered_circuit_needs_two_exchanges(n, world):
    """Staged execution: the 40-layer brick-work circuit needs 2 exchanges on 2, 4 and 8 ranks, and the
    final qubit map is a permutation."""
    import quant_iron_b200 as qi
    from quant_iron_b200 import sharded, workloads as w
    c = w.build_circuit(qi, n, w.random_layered_circuit(n, 40))
    pl = sharded.plan(n, world, c)
    assert pl["exchanges"] <= 3, pl["exchanges"]
    assert sorted(pl["final_layout"]) == list(range(n))
    assert sharded.plan(n, 1, c)["exchanges"] == 0


def test_circuit_lowering_splits_runs_by_gate_kind():
    """Circuit._lower (host only): operator gates -> one qi_apply_circuit run, consecutive PauliTimeEvolution gates ->
    one qi_apply_pauli_exp_sequence run, Measurement / PauliString gates alone; parametric gates are resolved at
    every lowering (never cached) and join the operator run."""
    import quant_iron_b200 as qi
    P = qi.Pauli
    zz = qi.PauliString.new(0.5).with_op(0, P.Z).with_op(1, P.Z)
    xx = qi.PauliString.new(0.25).with_op(1, P.X).with_op(2, P.X)
    prm = qi.Parameter.new([0.3])
    c = (qi.CircuitBuilder(3).h_gate(0).cnot_gate(1, 0)                                  # run 0: ops (2 records)
         .pauli_time_evolution_gate(zz, 0.1).pauli_time_evolution_gate(xx, 0.2)          # run 1: evol (2 terms)
         .rx_gate(2, 0.4).parametric_ry_gate(1, prm).z_gate(0)                           # run 2: ops (3 records)
         .measure_gate(qi.MeasurementBasis.Computational, [0])                           # run 3: gate (index 7)
         .pauli_string_gate(xx)                                                          # run 4: gate (index 8)
         .t_gate(2).build())                                                             # run 5: ops (1 record)
    runs = c._lower()
    assert [r[0] for r in runs] == ["ops", "evol", "ops", "gate", "gate", "ops"]
    assert [r[2] for r in runs if r[0] == "ops"] == [2, 3, 1]
    assert runs[1][2] == 2 and list(runs[1][4])[:4] == [0.0, -0.1, 0.0, -0.2]           # factors = (0, -t) per gate
    assert runs[3][2] == 7 and runs[4][2] == 8                                           # gate index = seed offset
    ry = runs[2][1][1]
    assert ry.kind == 12 and abs(ry.params[0] - 0.3) < 1e-15 and ry.targets[0] == 1      # QI_GATE_RY with the parameter
    prm.set([0.9])
    assert abs(c._lower()[2][1][1].params[0] - 0.9) < 1e-15                              # re-resolved, not cached
    plain = qi.CircuitBuilder(2).h_gate(0).cnot_gate(1, 0).build()
    assert plain._lower() is plain._lower()                                              # cached when nothing is parametric
