#!/usr/bin/env bash
set -u
OUT=gpurun_out/r02h
mkdir -p "$OUT"
timeout 900 python bench.py > "$OUT/bench_full.json" 2> "$OUT/bench_full.err"
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > "$OUT/bench_reference.json" 2> "$OUT/bench_reference.err"
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "tile_from_11q and (fuzz_random_gate_lists or random_layered)" > "$OUT/sanitizer_racecheck_tile.log" 2>&1
echo "exit $?" >> "$OUT/sanitizer_racecheck_tile.log"
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "tile_from_11q and (fuzz_random_gate_lists or qft_matches)" > "$OUT/sanitizer_memcheck_tile.log" 2>&1
echo "exit $?" >> "$OUT/sanitizer_memcheck_tile.log"
ls -la "$OUT"
