#!/bin/bash
# round 2: data-parallel step over peer memory vs NCCL (run with gpurun --gpus N); N from $1 (default 2)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2dp_topo.txt 2>&1
timeout 600 python -m pytest tests/test_dp.py -m gpu -q -x > gpurun_out/r2dp_pytest.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/r2dp_pytest.log
for mode in 1 0; do
  VAEASSOC_DP_PEER=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 100 --warmup 10 --no-parity > gpurun_out/r2dp_n${N}_peer$mode.json 2> gpurun_out/r2dp_n${N}_peer$mode.err
  echo "bench N=$N peer=$mode exit $?"; tail -3 gpurun_out/r2dp_n${N}_peer$mode.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2dp_n${N}_peer$mode.json").read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","n_gpus","dp_mode","gpu_launches")}, d["e2e"]["value"])
except Exception as e: print("no json", e)
PY
done
