#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dp.py -m gpu -q -x -s > gpurun_out/r2dp2_pytest.log 2>&1; echo "pytest exit $?"; grep -v "^\s*$" gpurun_out/r2dp2_pytest.log | tail -12
for mode in nvls ipc; do
  if [ $mode = nvls ]; then E="VAEASSOC_DP_SYMMETRIC=1"; else E="VAEASSOC_DP_SYMMETRIC=0"; fi
  env $E VAEASSOC_PEER_TIMELINE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus $N --steps 200 --warmup 10 --quick 2>&1 | grep -v "^\*\*\*\|OMP_NUM\|^$" | tail -4
done
