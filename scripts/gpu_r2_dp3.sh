#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dp.py -m gpu -q -x > gpurun_out/r2dp3_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2dp3_pytest.log
for mode in ipc nvls nccl; do
  case $mode in nvls) E="VAEASSOC_DP_SYMMETRIC=1";; ipc) E="VAEASSOC_DP_SYMMETRIC=0";; nccl) E="VAEASSOC_DP_PEER=0";; esac
  env $E VAEASSOC_PEER_TIMELINE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus $N --steps 200 --warmup 10 --quick > gpurun_out/r2dp3_n${N}_$mode.json 2> gpurun_out/r2dp3_n${N}_$mode.err
  echo "== $mode: $(cat gpurun_out/r2dp3_n${N}_$mode.json)"; grep "peer timeline" gpurun_out/r2dp3_n${N}_$mode.err | sort | head -2
done
