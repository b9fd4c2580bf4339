#!/bin/bash
# N-rank diagnostics of the data-parallel step: two-bucket overlapped schedule vs one all-reduce, NCCL channel limits
N=${1:-8}
run() { timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --steps 200 --warmup 20 --quick; }
echo "two buckets:";        run 29541 2> gpurun_out/dp_a.err
echo "single:";             VAEASSOC_DP_SINGLE=1 run 29542 2> gpurun_out/dp_b.err
echo "single, 4 channels:"; VAEASSOC_DP_SINGLE=1 NCCL_MAX_NCHANNELS=4 run 29543 2> gpurun_out/dp_c.err
echo "single, info:";       VAEASSOC_DP_SINGLE=1 NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL,TUNING run 29544 2> gpurun_out/dp_d.err
grep -i "algo\|channels\|nvls\|proto" gpurun_out/dp_d.err | head -40 > gpurun_out/dp_nccl_info.txt
