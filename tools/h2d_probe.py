"""Aggregate pinned host -> device copy bandwidth of the box with N ranks copying at once (names the limiter of the e2e leg
at N = 8).  Launch like bench.py: python -m torch.distributed.run --nproc-per-node N tools/h2d_probe.py"""
import json
import os
import time

import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nbytes, reps = 100_663_296, 60
    h = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(2)]
    d = [torch.empty(nbytes, dtype=torch.uint8, device="cuda") for _ in range(2)]
    for i in range(4):
        d[i & 1].copy_(h[i & 1], non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(reps):
        d[i & 1].copy_(h[i & 1], non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"probe": "pinned H2D, all ranks at once", "n_gpus": world, "bytes_per_copy": nbytes, "copies": reps,
                          "per_gpu_GBps": nbytes * reps / float(t.item()) / 1e9, "aggregate_GBps": world * nbytes * reps / float(t.item()) / 1e9}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
