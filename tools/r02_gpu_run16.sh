#!/usr/bin/env bash
set -u
OUT=gpurun_out/r02p
mkdir -p "$OUT"
timeout 900 python tools/exp_single_jit.py 30 > "$OUT/single_jit.txt" 2>&1
S="--skip-cpu --skip-extras --skip-e2e --steps 3"
timeout 600 python bench.py $S --opt jit_smem_kb=56 > "$OUT/bench_smem56.json" 2> "$OUT/bench_smem56.err"
timeout 600 python bench.py $S > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"
ls -la "$OUT"
