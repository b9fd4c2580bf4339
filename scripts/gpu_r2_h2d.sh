#!/bin/bash
# round 2: bare pinned H2D ceiling at 1 / 2 / 4 / 8 ranks on one 8-GPU box (scripts/h2d_probe.py) -> gpurun_out/r2_h2d_probe.jsonl
mkdir -p gpurun_out
: > gpurun_out/r2_h2d_probe.jsonl
for N in 1 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 scripts/h2d_probe.py 2>/dev/null | grep ranks >> gpurun_out/r2_h2d_probe.jsonl
done
cat gpurun_out/r2_h2d_probe.jsonl
