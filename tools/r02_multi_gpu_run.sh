#!/usr/bin/env bash
# Round-2 multi-GPU run (N = 2, 4 or 8 GPUs of one box):  gpurun --gpus N --timeout 1500 -- 'bash tools/r02_multi_gpu_run.sh N [cfg5_local_qubits]'
# sharded parity against the oracle (incl. the tile executor and its JIT modules on shards), the weak-scaling bench line with
# config 5 (QFT on N x 2^cfg5 amplitudes) and the sharded-vs-single-GPU probe parity in its extras, the NCCL half-shard swap A/B.
set -u
N=${1:-8}
Q5=${2:-33}
OUT=gpurun_out/r02_multi_$N
mkdir -p "$OUT"
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $RUN --master-port 29611 tests/sharded_worker.py --big 24 > "$OUT/sharded_parity.log" 2>&1
echo "parity exit $?" >> "$OUT/sharded_parity.log"
timeout 1500 $RUN --master-port 29612 bench.py --gpus $N --config5-local-qubits $Q5 > "$OUT/bench.json" 2> "$OUT/bench.err"
if [ "$N" -le 2 ]; then
  timeout 600 $RUN --master-port 29613 bench.py --gpus $N --skip-e2e --skip-extras --skip-cpu --opt jit=0 > "$OUT/bench_nojit.json" 2> "$OUT/bench_nojit.err"
fi
timeout 600 $RUN --master-port 29614 tools/nccl_halfshard_ab.py 30 > "$OUT/nccl_halfshard_ab.json" 2> "$OUT/nccl_halfshard_ab.err"
nvidia-smi topo -m > "$OUT/topo.txt" 2>&1
ls -la "$OUT"
