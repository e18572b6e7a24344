#!/usr/bin/env bash
# sharded parity only (N GPUs): gpurun --gpus N -- 'bash tools/r02_multi_parity_only.sh N'
set -u
N=${1:-4}
OUT=gpurun_out/r02_multi_$N
mkdir -p "$OUT"
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $RUN --master-port 29611 tests/sharded_worker.py --big 24 > "$OUT/sharded_parity.log" 2>&1
echo "parity exit $?" >> "$OUT/sharded_parity.log"
timeout 900 $RUN --master-port 29612 bench.py --gpus $N --skip-e2e --skip-cpu --config5-local-qubits 31 > "$OUT/bench.json" 2> "$OUT/bench.err"
ls -la "$OUT"
