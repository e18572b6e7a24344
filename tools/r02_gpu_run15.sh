#!/usr/bin/env bash
# round-2 record run on one B200: full GPU test tier, full bench (extras, e2e, cpu baseline), reference arm, sanitizers on the
# JIT / tile paths, ncu launch list of the bench command and one full capture of a module
set -u
OUT=gpurun_out/r02final2
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > "$OUT/gpu.csv" 2>&1
timeout 1500 python -m pytest tests -q -m gpu > "$OUT/pytest_gpu.log" 2>&1
echo "exit $?" >> "$OUT/pytest_gpu.log"
timeout 1200 python bench.py > "$OUT/bench_full.json" 2> "$OUT/bench_full.err"
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > "$OUT/bench_reference.json" 2> "$OUT/bench_reference.err"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches_bench.csv" python bench.py --steps 2 --warmup 3 --skip-cpu --skip-extras --skip-e2e > "$OUT/ncu_launches.log" 2>&1
timeout 600 ncu --set full --clock-control none -k regex:qi_tile_jit -s 40 -c 3 -o "$OUT/jit_full" python bench.py --steps 1 --skip-cpu --skip-extras --skip-e2e > "$OUT/ncu_full.log" 2>&1
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_tile_jit.py -q -m gpu -k "fuzzed or variants or qft" > "$OUT/sanitizer_memcheck_jit.log" 2>&1
echo "exit $?" >> "$OUT/sanitizer_memcheck_jit.log"
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_tile_jit.py -q -m gpu -k "layered or variants" > "$OUT/sanitizer_racecheck_jit.log" 2>&1
echo "exit $?" >> "$OUT/sanitizer_racecheck_jit.log"
python __graft_entry__.py --smoke > "$OUT/smoke.log" 2>&1
ls -la "$OUT"
