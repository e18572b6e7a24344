"""Multi-rank tests of the sharded state.  Host logic on CPU over gloo (world_size 2), the engine
itself on >= 2 GPUs over torchrun (skipped on a single-GPU box)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "sharded_worker.py")


def _torchrun(nproc, port, extra):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), WORKER] + extra
    return subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)


def test_host_logic_gloo_world2():
    r = _torchrun(2, 29531, ["--cpu"])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "PASS cpu host-logic world=2" in r.stdout


def test_planner_matches_survey_counts():
    """in QFT-36 on 8 GPUs every controlled phase is communication-free; only the three
    Hadamards on global qubits need their qubits brought in, which the engine does in ONE 3-qubit
    all-to-all exchange (the final swaps are relabelled)."""
    import quant_iron_b200 as qi
    from quant_iron_b200 import sharded, workloads as w
    c = w.build_circuit(qi, 36, w.qft_specs(36))
    pl = sharded.plan(36, 8, c)
    assert pl["exchanges"] == 1
    assert pl["comm_free_global_gates"] >= 99
    assert sorted(pl["final_layout"]) == list(range(36))
    assert sharded.plan(33, 1, w.build_circuit(qi, 33, w.qft_specs(33)))["exchanges"] == 0


def test_pauli_planner_stages_trotter_steps():
    """Sharded Trotter evolution is staged around the exchanges with the exact Pauli commutation test: the
    Heisenberg chain needs about ONE exchange per Trotter step (one per term that flips a rank-bit qubit without
    staging: 6-7 per step), and an unsharded state needs none."""
    import quant_iron_b200 as qi
    from quant_iron_b200 import sharded
    h = qi.heisenberg_1d(33, 1.0, 2.0, 3.0, 0.5, 0.1)
    p1, p10 = sharded.plan_pauli(33, 8, h, 1), sharded.plan_pauli(33, 8, h, 10)
    assert p1["exchanges"] <= 2 and p10["exchanges"] <= 12, (p1, p10)
    assert p10["stages"] == p10["exchanges"] + 1
    assert sharded.plan_pauli(33, 1, h, 3) == {"exchanges": 0, "stages": 1}
    # a purely diagonal Hamiltonian never communicates
    assert sharded.plan_pauli(20, 8, qi.ising_1d_uniform(20, 1.0, 2.0, 0.1), 5)["exchanges"] == 0


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_engine_vs_oracle(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    r = _torchrun(world, 29540 + world, ["--big", "24"])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "ALL PASS" in r.stdout
