#!/usr/bin/env bash
set -u
OUT=gpurun_out/r02x
mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_tile_jit.py -q -m gpu > "$OUT/pytest.log" 2>&1
echo "exit $?" >> "$OUT/pytest.log"
S="--skip-cpu --skip-extras --skip-e2e --steps 3"
timeout 600 python bench.py $S --opt tile_restore=1 > "$OUT/bench_restore.json" 2> "$OUT/bench_restore.err"
timeout 600 python bench.py $S > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"
timeout 600 python bench.py $S --opt tile_restore=1 > "$OUT/bench_restore_2.json" 2> "$OUT/bench_restore_2.err"
ls -la "$OUT"
