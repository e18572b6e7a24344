#!/usr/bin/env bash
# Multi-GPU counterpart of first_gpu_run.sh (N = 2, 4 or 8 GPUs of one box):
#   gpurun --gpus 8 --timeout 1200 -- 'bash tools/first_multi_gpu_run.sh 8'
# sharded parity against the oracle, the weak-scaling bench line, and the same with the lean forms.
set -u
N=${1:-8}
OUT=gpurun_out/first_multi_$N
mkdir -p "$OUT"
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$RUN --master-port 29611 tests/sharded_worker.py --big 24 > "$OUT/sharded_parity.log" 2>&1
echo "parity exit $?" >> "$OUT/sharded_parity.log"
$RUN --master-port 29612 bench.py --gpus $N > "$OUT/bench.json" 2> "$OUT/bench.err"
$RUN --master-port 29613 bench.py --gpus $N --skip-e2e --opt lean=1 > "$OUT/bench_lean.json" 2> "$OUT/bench_lean.err"
$RUN --master-port 29614 bench.py --gpus $N --skip-e2e --opt late_tables=0 > "$OUT/bench_early_tables.json" 2> "$OUT/bench_early_tables.err"
python bench.py --impl reference --gpus $N --steps 2 --warmup 1 > "$OUT/bench_reference.json" 2> "$OUT/bench_reference.err"
ls -la "$OUT"
