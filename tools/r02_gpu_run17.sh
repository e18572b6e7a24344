#!/usr/bin/env bash
set -u
OUT=gpurun_out/r02q
mkdir -p "$OUT"
timeout 1500 python -m pytest tests -q -m "gpu and not large" > "$OUT/pytest_gpu.log" 2>&1
echo "exit $?" >> "$OUT/pytest_gpu.log"
timeout 1200 python bench.py > "$OUT/bench_full.json" 2> "$OUT/bench_full.err"
ls -la "$OUT"
