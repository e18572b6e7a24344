#!/usr/bin/env bash
# JIT modules: L2 prefetch A/B, fixed bit-identity tests, QFT 33 per executor
set -u
OUT=gpurun_out/r02m
mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_tile_jit.py -q -m gpu > "$OUT/pytest_jit.log" 2>&1
echo "exit $?" >> "$OUT/pytest_jit.log"
S="--skip-cpu --skip-extras --skip-e2e --steps 3"
timeout 600 python bench.py $S --opt jit_prefetch=1 > "$OUT/bench_prefetch.json" 2> "$OUT/bench_prefetch.err"
timeout 600 python bench.py $S > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"
timeout 600 python bench.py $S --opt jit_prefetch=1 --opt jit_ctas=3 > "$OUT/bench_prefetch_ctas3.json" 2> "$OUT/bench_prefetch_ctas3.err"
timeout 900 python tools/prof_qft.py 33 2 > "$OUT/qft33.txt" 2>&1
ls -la "$OUT"
