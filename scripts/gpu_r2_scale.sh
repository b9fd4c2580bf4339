#!/bin/bash
# round 2: the driver's scaling run in small -- bench.py at N = 4 and N = 8 on one 8-GPU box (default data-parallel mode)
mkdir -p gpurun_out
for N in 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N \
    bench.py --gpus $N --steps 100 --warmup 10 --no-cpu-baseline --no-secondary > gpurun_out/r2scale_n$N.json 2> gpurun_out/r2scale_n$N.err
  echo "N=$N exit $?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2scale_n$N.json").read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","n_gpus","dp_mode")}, "e2e", d.get("e2e",{}).get("value"), "dd", d.get("e2e_device_dataset",{}).get("value"), "parity", (d.get("parity") or {}).get("worst_grad_l2"))
except Exception as e: print("no json", e)
PY
done
