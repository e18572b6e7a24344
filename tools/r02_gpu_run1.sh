#!/usr/bin/env bash
# Round 2, first GPU call (one B200): design-deciding microbenchmarks, the parity suite, the A/B of the options round 1
# built blind, full ncu captures of the Pauli kernels, and a compute-sanitizer pass.  gpurun --timeout 900 -- 'bash tools/r02_gpu_run1.sh'
set -u
OUT=gpurun_out/r02a
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw --format=csv > "$OUT/gpu.csv" 2>&1
nproc > "$OUT/host.txt"; grep -m1 "model name" /proc/cpuinfo >> "$OUT/host.txt"; free -g >> "$OUT/host.txt"
timeout 120 tools/micro/tile_proto 30 > "$OUT/tile_proto.txt" 2>&1
timeout 60 tools/micro/fp64_peak > "$OUT/fp64_peak.txt" 2>&1
timeout 300 python -m pytest tests -q -m gpu -x > "$OUT/pytest_gpu.log" 2>&1
echo "pytest exit $?" >> "$OUT/pytest_gpu.log"
B="python bench.py --steps 5 --warmup 3 --skip-cpu --skip-extras --skip-e2e"
for rep in 1 2; do
  timeout 200 $B > "$OUT/bench_default_$rep.json" 2> "$OUT/bench_default_$rep.err"
  timeout 200 $B --opt lean=1 > "$OUT/bench_lean_$rep.json" 2> "$OUT/bench_lean_$rep.err"
done
timeout 200 $B --opt lean=1 --opt prefetch=1 > "$OUT/bench_lean_prefetch.json" 2> "$OUT/bench_lean_prefetch.err"
timeout 200 $B --opt lean=1 --opt window_regs=5 > "$OUT/bench_lean_r5.json" 2> "$OUT/bench_lean_r5.err"
timeout 200 $B --opt lean=1 --opt tma=1 > "$OUT/bench_lean_tma.json" 2> "$OUT/bench_lean_tma.err"
timeout 200 $B --opt late_tables=0 > "$OUT/bench_early_tables.json" 2> "$OUT/bench_early_tables.err"
# full captures of the Pauli kernels (VERDICT weak #3): 24 sites (config 3) and 28 sites
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_pauli_window -s 4 -c 2 -o "$OUT/pauli_window_24" python tools/prof_trotter.py 24 2 > "$OUT/ncu_pauli.log" 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_pauli_window -s 2 -c 2 -o "$OUT/pauli_window_28" python tools/prof_trotter.py 28 1 >> "$OUT/ncu_pauli.log" 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_pauli_expect_window -s 1 -c 2 -o "$OUT/pauli_expect_24" python tools/prof_expect.py 24 2 >> "$OUT/ncu_pauli.log" 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_pauli_expect_window -s 1 -c 2 -o "$OUT/pauli_expect_28" python tools/prof_expect.py 28 2 >> "$OUT/ncu_pauli.log" 2>&1
# memcheck: smoke() and the fused-executor fuzz test (n <= 13)
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/sanitizer_memcheck_smoke.log" 2>&1
echo "exit $?" >> "$OUT/sanitizer_memcheck_smoke.log"
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "fuzz_random_gate_lists or tma_prefetched" > "$OUT/sanitizer_memcheck_fuzz.log" 2>&1
echo "exit $?" >> "$OUT/sanitizer_memcheck_fuzz.log"
ls -la "$OUT"
