#!/usr/bin/env bash
# Round 2, third GPU call: k_tile with in-place updates and the CX/H -> CZ rewrite; A/B of absorb / cz_rewrite / lean.
set -u
OUT=gpurun_out/r02c
mkdir -p "$OUT"
timeout 600 python -m pytest tests -q -m gpu -x > "$OUT/pytest_gpu.log" 2>&1
echo "pytest exit $?" >> "$OUT/pytest_gpu.log"
B="python bench.py --steps 5 --warmup 3 --skip-cpu --skip-extras --skip-e2e"
run() { name=$1; shift; timeout 200 $B "$@" > "$OUT/bench_$name.json" 2> "$OUT/bench_$name.err"; }
run default
run absorb0 --opt absorb=0
run cz0 --opt cz_rewrite=0
run lean --opt lean=1
run absorb0_lean --opt absorb=0 --opt lean=1
run default_2
run window --opt tile=0
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches_absorb0.csv" $B --steps 1 --opt absorb=0 > "$OUT/ncu_bench.log" 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_tile -s 10 -c 2 -o "$OUT/tile_full_absorb0" $B --steps 1 --opt absorb=0 > "$OUT/ncu_full.log" 2>&1
ls -la "$OUT"
