#!/usr/bin/env bash
# JIT modules: staged loads (bulk async copies into shared memory) and late L2 prefetch, A/B on one box
set -u
OUT=gpurun_out/r02n
mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_tile_jit.py -q -m gpu > "$OUT/pytest_jit.log" 2>&1
echo "exit $?" >> "$OUT/pytest_jit.log"
S="--skip-cpu --skip-extras --skip-e2e --steps 3"
timeout 600 python bench.py $S --opt jit_stage=1 > "$OUT/bench_stage.json" 2> "$OUT/bench_stage.err"
timeout 600 python bench.py $S > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"
timeout 600 python bench.py $S --opt jit_prefetch=1 > "$OUT/bench_prefetch_late.json" 2> "$OUT/bench_prefetch_late.err"
timeout 600 ncu --set full --clock-control none -k regex:qi_tile_jit -s 40 -c 2 -o "$OUT/jit_stage_full" python bench.py $S --steps 1 --opt jit_stage=1 > "$OUT/ncu_stage.log" 2>&1
ls -la "$OUT"
