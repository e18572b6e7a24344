#!/usr/bin/env bash
set -u
OUT=gpurun_out/r02g
mkdir -p "$OUT"
timeout 600 python -m pytest tests -q -m gpu -x > "$OUT/pytest_gpu.log" 2>&1
echo "pytest exit $?" >> "$OUT/pytest_gpu.log"
B="python bench.py --steps 5 --warmup 3 --skip-cpu --skip-extras --skip-e2e"
run() { name=$1; shift; timeout 200 $B "$@" > "$OUT/bench_$name.json" 2> "$OUT/bench_$name.err"; }
run default
run cz0 --opt cz_rewrite=0
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_tile -s 19 -c 1 -o "$OUT/tile_full" $B --steps 1 > "$OUT/ncu_full.log" 2>&1
ls -la "$OUT"
