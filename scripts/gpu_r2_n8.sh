#!/bin/bash
# round 2: 8-GPU pass -- the data-parallel step over peer memory (the default) at N = 8 on the reference config (8192
# pairs per GPU = global 65 536, BASELINE configs[2]) and on the scaled config (16 384 per GPU = global 131 072, configs[4])
N=${1:-8}
mkdir -p gpurun_out
run() {  # name, env, extra args
  env $2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 \
    bench.py --gpus $N $3 > gpurun_out/r2n${N}_$1.json 2> gpurun_out/r2n${N}_$1.err
  echo "$1 exit $?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2n${N}_$1.json").read().strip().splitlines()[-1])
    print("$1", {k:d.get(k) for k in ("value","ms_per_step","n_gpus","dp_mode")}, "e2e", d.get("e2e",{}).get("value"), "dd", d.get("e2e_device_dataset",{}).get("value"), "parity", (d.get("parity") or {}).get("worst_grad_l2"))
except Exception as e: print("$1 no json", e)
PY
}
run peer "X=1" "--steps 100 --warmup 10 --no-cpu-baseline --no-secondary"
run scaled_peer "X=1" "--config scaled --steps 20 --warmup 5 --no-cpu-baseline --no-secondary --no-parity"
tail -3 gpurun_out/r2n${N}_peer.err
