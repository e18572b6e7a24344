#!/usr/bin/env bash
set -u
OUT=gpurun_out/r02s
mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_tile_jit.py tests/test_gpu_parity.py -q -m gpu -x -k "jit or layered or qft or fuzz" > "$OUT/pytest.log" 2>&1
echo "exit $?" >> "$OUT/pytest.log"
S="--skip-cpu --skip-extras --skip-e2e --steps 3"
timeout 600 python bench.py $S > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"
timeout 600 python bench.py $S --opt jit_ctas=5 > "$OUT/bench_ctas5.json" 2> "$OUT/bench_ctas5.err"
timeout 600 python bench.py $S --opt jit_ctas=6 > "$OUT/bench_ctas6.json" 2> "$OUT/bench_ctas6.err"
timeout 600 python bench.py $S > "$OUT/bench_default_2.json" 2> "$OUT/bench_default_2.err"
ls -la "$OUT"
