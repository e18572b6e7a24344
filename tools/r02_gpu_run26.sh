#!/usr/bin/env bash
set -u
OUT=gpurun_out/r02ze
mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_tile_jit.py tests/test_gpu_parity.py -q -m gpu -x > "$OUT/pytest.log" 2>&1
echo "exit $?" >> "$OUT/pytest.log"
timeout 900 python tools/prof_qft_restore.py 33 > "$OUT/qft33_cycle.txt" 2>&1
S="--skip-cpu --skip-extras --skip-e2e --steps 3"
timeout 600 python bench.py $S > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"
timeout 900 python -m pytest tests/test_gpu_parity_large.py -q -m gpu -k "cp_ladders" > "$OUT/pytest_large_cp.log" 2>&1
echo "exit $?" >> "$OUT/pytest_large_cp.log"
ls -la "$OUT"
