#!/usr/bin/env bash
# Round 2, GPU call 6: all-in-place op set of k_tile
set -u
OUT=gpurun_out/r02f
mkdir -p "$OUT"
timeout 600 python -m pytest tests -q -m gpu -x > "$OUT/pytest_gpu.log" 2>&1
echo "pytest exit $?" >> "$OUT/pytest_gpu.log"
B="python bench.py --steps 5 --warmup 3 --skip-cpu --skip-extras --skip-e2e"
run() { name=$1; shift; timeout 200 $B "$@" > "$OUT/bench_$name.json" 2> "$OUT/bench_$name.err"; }
run default
run cz0 --opt cz_rewrite=0
run noslide --opt tile_slide=0
run window --opt tile=0
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_tile -s 18 -c 2 -o "$OUT/tile_full" $B --steps 1 > "$OUT/ncu_full.log" 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches.csv" $B --steps 1 > "$OUT/ncu_bench.log" 2>&1
ls -la "$OUT"
