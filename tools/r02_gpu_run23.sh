#!/usr/bin/env bash
set -u
OUT=gpurun_out/r02zb
mkdir -p "$OUT"
timeout 900 python -m pytest tests -q -m "gpu and not large" -k "pauli or trotter or heisenberg or sumop or expect or golden or models" > "$OUT/pytest_pauli.log" 2>&1
echo "exit $?" >> "$OUT/pytest_pauli.log"
timeout 600 python tools/prof_trotter.py 24 50 > "$OUT/trotter24.txt" 2>&1
timeout 600 python tools/prof_trotter.py 28 10 > "$OUT/trotter28.txt" 2>&1
timeout 900 python -m pytest tests/test_gpu_parity_large.py -q -m gpu -k heisenberg > "$OUT/pytest_large_heis.log" 2>&1
echo "exit $?" >> "$OUT/pytest_large_heis.log"
ls -la "$OUT"
