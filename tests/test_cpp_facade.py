"""The compiled-language host side: include/quant_iron_b200.hpp (C++ facade over the C ABI) passes a
selection of the reference's known-answer tests on the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "test_facade")


def test_facade_compiles_on_cpu_box():
    import __graft_entry__ as g
    g.build()
    assert os.path.exists(BIN)


def test_facade_host_logic_on_cpu():
    """Model builders, builder validation, QFT gate counts and the unitarity check of the C++ facade: no device."""
    host = os.path.join(ROOT, "tests", "cpp", "test_facade_host")
    if not os.path.exists(host):
        import __graft_entry__ as g
        g.build()
    r = subprocess.run([host], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ALL PASS" in r.stdout


@pytest.mark.gpu
def test_facade_known_answers_on_gpu():
    if not os.path.exists(BIN):
        import __graft_entry__ as g
        g.build()
    r = subprocess.run([BIN], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ALL PASS" in r.stdout


@pytest.mark.gpu
@pytest.mark.unproven
def test_facade_execute_host_on_gpu():
    """Circuit::execute_host_ (qi_execute_host, pipelined) through the C++ facade: same state as execute()."""
    if not os.path.exists(BIN):
        import __graft_entry__ as g
        g.build()
    r = subprocess.run([BIN, "--execute-host"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ALL PASS" in r.stdout


@pytest.mark.gpu
@pytest.mark.unproven
def test_facade_wider_surface_on_gpu():
    """State gate families, Bell / Hartree-Fock constructors, metrics, SumOp::apply, circuits with measurement and Pauli
    gates, measure / measure_n with custom bases through the C++ facade."""
    if not os.path.exists(BIN):
        import __graft_entry__ as g
        g.build()
    r = subprocess.run([BIN, "--surface"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ALL PASS" in r.stdout
