"""Bare pinned host -> device copy ceiling per number of ranks (the bound of bench.py's `e2e`: one 30.5 MB batch per step
and rank from pinned host memory).  Under torchrun: every rank copies `--mb` megabytes `--iters` times from two pinned
buffers on its own stream, all ranks at once; prints per-rank and aggregate GB/s (device-timed, max over ranks)."""
import argparse
import json
import os

import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--mb", type=float, default=30.507008)
ap.add_argument("--iters", type=int, default=200)
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(args.mb * 1e6) // 4
host = [torch.empty(n, dtype=torch.float32).pin_memory() for _ in range(2)]
for h in host:
    h.normal_()
dev = [torch.empty(n, dtype=torch.float32, device="cuda") for _ in range(2)]
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for i in range(10):
        dev[i & 1].copy_(host[i & 1], non_blocking=True)
s.synchronize()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(s):
    e0.record()
    for i in range(args.iters):
        dev[i & 1].copy_(host[i & 1], non_blocking=True)
    e1.record()
s.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
gbs = n * 4 * args.iters / (ms.item() * 1e-3) / 1e9
if rank == 0:
    print(json.dumps(dict(ranks=world, mb_per_copy=args.mb, iters=args.iters, gbs_per_rank=gbs, gbs_aggregate=gbs * world,
                          pairs_per_s_ceiling=gbs * world * 1e9 / 3724.0)))
if world > 1:
    dist.destroy_process_group()
