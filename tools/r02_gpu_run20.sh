#!/usr/bin/env bash
set -u
OUT=gpurun_out/r02v
mkdir -p "$OUT"
S="--skip-cpu --skip-extras --skip-e2e --steps 3"
timeout 600 python bench.py $S > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"
for ns in 10000 20000 40000; do
  timeout 600 python bench.py $S --opt jit_prefetch=$ns > "$OUT/bench_stagger_$ns.json" 2> "$OUT/bench_stagger_$ns.err"
done
timeout 600 python bench.py $S > "$OUT/bench_default_2.json" 2> "$OUT/bench_default_2.err"
ls -la "$OUT"
