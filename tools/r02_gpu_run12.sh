#!/usr/bin/env bash
# JIT tile modules, third run: unit-form (lean) pair ops, per-group W tables, QFT with the layout period warmed, large parity
set -u
OUT=gpurun_out/r02l
mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_tile_jit.py -q -m gpu > "$OUT/pytest_jit.log" 2>&1
echo "exit $?" >> "$OUT/pytest_jit.log"
S="--skip-cpu --skip-extras --skip-e2e --steps 3"
timeout 600 python bench.py $S > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"
timeout 600 python bench.py $S --opt tile_lean=0 > "$OUT/bench_nolean.json" 2> "$OUT/bench_nolean.err"
timeout 600 python bench.py $S --opt jit_groups=2 > "$OUT/bench_g2.json" 2> "$OUT/bench_g2.err"
timeout 600 python bench.py $S --opt jit=0 > "$OUT/bench_nojit.json" 2> "$OUT/bench_nojit.err"
timeout 600 python tools/prof_qft.py 30 4 > "$OUT/qft30.txt" 2>&1
timeout 600 ncu --set full --clock-control none -k regex:qi_tile_jit -s 40 -c 2 -o "$OUT/jit_full" python bench.py $S --steps 1 > "$OUT/ncu_full.log" 2>&1
timeout 1200 python -m pytest tests/test_gpu_parity_large.py -q -m gpu > "$OUT/pytest_large.log" 2>&1
echo "exit $?" >> "$OUT/pytest_large.log"
timeout 900 python -m pytest tests -q -m "gpu and not large" -x > "$OUT/pytest_gpu.log" 2>&1
echo "exit $?" >> "$OUT/pytest_gpu.log"
ls -la "$OUT"
