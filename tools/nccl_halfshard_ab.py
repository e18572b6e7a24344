"""A/B baseline for the shard exchange (north_star (5) words it as "an NCCL sendrecv swap of half-shards over NVLink").

Every rank trades HALF of a 2^q-amplitude f64 shard with its XOR partner (rank ^ 1) through ncclSend/ncclRecv
(torch.distributed.batch_isend_irecv on the NCCL backend).  NCCL cannot swap in place: the received half lands in a staging
buffer of the same size (at 33 local qubits that is 64 GiB next to a 128 GiB shard -- it does not fit, which is why the engine
swaps in place through peer memory, csrc/shard.cu k_exchange) and is then copied over the half that was sent.  Reported:
GB/s per GPU per direction for (a) the NCCL transfer alone and (b) transfer + the copy back, to be read next to
`comm.nvlink_gbs_per_gpu_per_direction` of `bench.py --gpus N` (the in-place peer-memory kernel).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/nccl_halfshard_ab.py [qubits]
"""
import json
import os
import sys

import torch
import torch.distributed as dist


def main():
    q = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    half = 1 << (q - 1)                                   # amplitudes in a half shard
    shard = torch.zeros(2 * half, dtype=torch.complex128, device="cuda")
    shard[:] = rank + 1
    stage = torch.empty(half, dtype=torch.complex128, device="cuda")
    partner = rank ^ 1
    mine = shard[half:] if rank < partner else shard[:half]      # the half that leaves

    def swap(copy_back):
        ops = [dist.P2POp(dist.isend, mine, partner), dist.P2POp(dist.irecv, stage, partner)]
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        if copy_back:
            mine.copy_(stage)

    out = {}
    for name, cb in (("nccl_sendrecv_only", False), ("nccl_sendrecv_plus_copy_back", True)):
        for _ in range(2):
            swap(cb)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            swap(cb)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device="cuda")
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        out[name] = {"ms": float(ms.item()), "gbs_per_gpu_per_direction": 16.0 * half / (float(ms.item()) * 1e-3) / 1e9}
    ok = bool((mine == partner + 1).all().item())
    if rank == 0:
        print(json.dumps({"what": "NCCL half-shard swap with the XOR partner", "world": world, "local_qubits": q,
                          "bytes_per_direction": 16 * half, "swapped_values_ok": ok, **out}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
