#!/usr/bin/env bash
# JIT tile modules, second run: layout-stable passes (no slide), tile groups per CTA (instruction-stream sharing), QFT, large parity
set -u
OUT=gpurun_out/r02k
mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_tile_jit.py -q -m gpu -x > "$OUT/pytest_jit.log" 2>&1
echo "exit $?" >> "$OUT/pytest_jit.log"
S="--skip-cpu --skip-extras --skip-e2e --steps 3"
for g in 2 1 4; do
  timeout 600 python bench.py $S --opt jit_groups=$g > "$OUT/bench_jit_g$g.json" 2> "$OUT/bench_jit_g$g.err"
done
timeout 600 python bench.py $S --opt jit_ctas=3 > "$OUT/bench_jit_ctas3.json" 2> "$OUT/bench_jit_ctas3.err"
timeout 600 python tools/prof_qft.py 30 3 > "$OUT/qft30.txt" 2>&1
timeout 600 ncu --set full --clock-control none -k regex:qi_tile_jit -s 40 -c 2 -o "$OUT/jit_full" python bench.py $S --steps 1 > "$OUT/ncu_full.log" 2>&1
timeout 1200 python -m pytest tests/test_gpu_parity_large.py -q -m gpu -x > "$OUT/pytest_large.log" 2>&1
echo "exit $?" >> "$OUT/pytest_large.log"
ls -la "$OUT"
