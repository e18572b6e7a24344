#!/usr/bin/env bash
set -u
OUT=gpurun_out/r02zc
mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_tile_jit.py tests/test_gpu_parity.py tests/test_golden.py -q -m gpu -x > "$OUT/pytest.log" 2>&1
echo "exit $?" >> "$OUT/pytest.log"
S="--skip-cpu --skip-extras --skip-e2e --steps 3"
timeout 600 python bench.py $S > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"
timeout 600 python bench.py $S > "$OUT/bench_default_2.json" 2> "$OUT/bench_default_2.err"
timeout 600 python tools/prof_qft.py 30 4 > "$OUT/qft30.txt" 2>&1
ls -la "$OUT"
