#!/usr/bin/env bash
# first run of the JIT tile modules: parity (bit-exact vs k_tile, vs oracle), A/B bench, ncu of one module
set -u
OUT=gpurun_out/r02j
mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_tile_jit.py -q -m gpu -x > "$OUT/pytest_jit.log" 2>&1
echo "exit $?" >> "$OUT/pytest_jit.log"
S="--skip-cpu --skip-extras --skip-e2e --steps 3"
timeout 600 python bench.py $S > "$OUT/bench_jit.json" 2> "$OUT/bench_jit.err"
timeout 600 python bench.py $S --opt jit=0 > "$OUT/bench_nojit.json" 2> "$OUT/bench_nojit.err"
timeout 600 python bench.py $S --opt jit_ctas=3 > "$OUT/bench_jit_ctas3.json" 2> "$OUT/bench_jit_ctas3.err"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:qi_tile_jit -s 20 -c 2 -o "$OUT/jit_full" python bench.py $S --steps 1 > "$OUT/ncu_full.log" 2>&1
timeout 900 python -m pytest tests -q -m gpu -x > "$OUT/pytest_gpu.log" 2>&1
echo "exit $?" >> "$OUT/pytest_gpu.log"
ls -la "$OUT"
