#!/usr/bin/env python
"""Host<->device copy bandwidth with all ranks copying at once (diagnostic for the e2e line of bench.py):
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/pcie_probe.py"""
import json
import os

import torch
import torch.distributed as dist

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 30
host = torch.empty(n, dtype=torch.uint8).pin_memory()
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
host.fill_(1)


def timed(fn, reps=4):
    fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return n / (t.item() * 1e-3) / 1e9


d2h = timed(lambda: host.copy_(dev, non_blocking=True))
h2d = timed(lambda: dev.copy_(host, non_blocking=True))
if rank == 0:
    print(json.dumps({"ranks": world, "d2h_GBps_per_rank_min": d2h, "h2d_GBps_per_rank_min": h2d, "d2h_aggregate": d2h * world,
                      "h2d_aggregate": h2d * world}))
if world > 1:
    dist.destroy_process_group()
