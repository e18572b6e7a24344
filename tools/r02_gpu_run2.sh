#!/usr/bin/env bash
# Round 2, second GPU call: first run of the CTA-tile executor (k_tile): parity suite, A/B against the warp-tile executor, launch list.
set -u
OUT=gpurun_out/r02b
mkdir -p "$OUT"
timeout 600 python -m pytest tests -q -m gpu -x > "$OUT/pytest_gpu.log" 2>&1
echo "pytest exit $?" >> "$OUT/pytest_gpu.log"
B="python bench.py --steps 5 --warmup 3 --skip-cpu --skip-extras --skip-e2e"
timeout 200 $B > "$OUT/bench_tile.json" 2> "$OUT/bench_tile.err"
timeout 200 $B --opt tile=0 > "$OUT/bench_window.json" 2> "$OUT/bench_window.err"
timeout 200 $B --opt lean=1 > "$OUT/bench_tile_lean.json" 2> "$OUT/bench_tile_lean.err"
timeout 200 $B > "$OUT/bench_tile_2.json" 2> "$OUT/bench_tile_2.err"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches_tile.csv" $B --steps 1 > "$OUT/ncu_bench.log" 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches_tile_lean.csv" $B --steps 1 --opt lean=1 > "$OUT/ncu_bench_lean.log" 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_tile -s 10 -c 3 -o "$OUT/tile_full" $B --steps 1 > "$OUT/ncu_full.log" 2>&1
ls -la "$OUT"
