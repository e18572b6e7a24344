#!/usr/bin/env bash
# First GPU call after a stretch of CPU-only work: validates everything that was built blind and collects the A/B
# numbers in ONE box (box-to-box variance is 3-5 %).  Run from the repo root:
#   gpurun --timeout 1500 -- 'bash tools/first_gpu_run.sh'
# Outputs land in gpurun_out/first/ (copy what should be judged into profiles/).
set -u
OUT=gpurun_out/first
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw --format=csv > "$OUT/gpu.csv" 2>&1

# 1. parity: established tests first, `unproven` ones (host pipeline, lean forms) last
python -m pytest tests -q -m gpu -x > "$OUT/pytest_gpu.log" 2>&1
echo "pytest exit $?" >> "$OUT/pytest_gpu.log"
# without -x, only the unproven ones, so that one failure does not hide the others
python -m pytest tests -q -m "gpu and unproven" > "$OUT/pytest_unproven.log" 2>&1

# 2. bench A/B on this box, alternating (default = late_tables on, lean off)
for rep in 1 2; do
  python bench.py --skip-cpu --skip-extras                       > "$OUT/bench_default_$rep.json"   2> "$OUT/bench_default_$rep.err"
  python bench.py --skip-cpu --skip-extras --skip-e2e --opt lean=1        > "$OUT/bench_lean_$rep.json"      2> "$OUT/bench_lean_$rep.err"
  python bench.py --skip-cpu --skip-extras --skip-e2e --opt late_tables=0 > "$OUT/bench_early_tables_$rep.json" 2> "$OUT/bench_early_tables_$rep.err"
done
python bench.py --skip-cpu --skip-extras --skip-e2e --opt lean=1 --opt window_regs=5 > "$OUT/bench_lean_r5.json" 2> "$OUT/bench_lean_r5.err"
python bench.py --skip-cpu --skip-extras --skip-e2e --opt lean=1 --opt tma=1         > "$OUT/bench_lean_tma.json" 2> "$OUT/bench_lean_tma.err"
python bench.py --skip-cpu --skip-extras --skip-e2e --opt lean=1 --opt prefetch=1    > "$OUT/bench_lean_prefetch.json" 2> "$OUT/bench_lean_prefetch.err"

# 3. the full default line (cpu baseline, single-gate table, e2e) and the reference arm
python bench.py > "$OUT/bench_full.json" 2> "$OUT/bench_full.err"
python bench.py --impl reference --steps 2 --warmup 1 > "$OUT/bench_reference.json" 2> "$OUT/bench_reference.err"

# 4. launch list of the bench command and one full capture of the dominant kernel (default and lean)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches_default.csv" \
    python bench.py --steps 2 --warmup 3 --skip-cpu --skip-extras --skip-e2e > "$OUT/ncu_bench.log" 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches_lean.csv" \
    python bench.py --steps 2 --warmup 3 --skip-cpu --skip-extras --skip-e2e --opt lean=1 >> "$OUT/ncu_bench.log" 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_window -s 130 -c 3 -o "$OUT/window_default" \
    python bench.py --steps 1 --warmup 3 --skip-cpu --skip-extras --skip-e2e >> "$OUT/ncu_bench.log" 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_window -s 130 -c 3 -o "$OUT/window_lean" \
    python bench.py --steps 1 --warmup 3 --skip-cpu --skip-extras --skip-e2e --opt lean=1 >> "$OUT/ncu_bench.log" 2>&1
ls -la "$OUT"
