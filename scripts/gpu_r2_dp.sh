#!/bin/bash
# round 2: data-parallel pass at N ranks with the one-launch step -- the 2-GPU tests, then bench lines in peer and NCCL mode
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dp.py -m gpu -q > gpurun_out/r2dp4_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2dp4_pytest.log
for mode in peer nccl; do
  case $mode in peer) E="VAEASSOC_DP_PEER=1";; nccl) E="VAEASSOC_DP_PEER=0";; esac
  env $E timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus $N --steps 100 --warmup 10 --no-cpu-baseline --no-secondary > gpurun_out/r2dp4_n${N}_$mode.json 2> gpurun_out/r2dp4_n${N}_$mode.err
  echo "== $mode exit $?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2dp4_n${N}_$mode.json").read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","n_gpus","dp_mode")}, "e2e", d.get("e2e",{}).get("value"), "parity", (d.get("parity") or {}).get("worst_grad_l2"))
except Exception as e: print("no json", e)
PY
  tail -2 gpurun_out/r2dp4_n${N}_$mode.err
done
