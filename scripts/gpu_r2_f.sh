#!/bin/bash
# round 2, pass F: one-launch form (loss fused into the output-layer epilogue, finalize task) -- parity tests, then bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2f_pytest.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/r2f_pytest.log
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench exit $?"; tail -3 gpurun_out/r2f_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2f_bench.json").read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","gpu_launches")}, "e2e", d["e2e"]["value"], "dd", d["e2e_device_dataset"]["value"])
    print("roofline", d["roofline"]["kernel"], d["roofline"]["frac"], "parity", d.get("parity"))
    for k in d["kernels"][:10]: print("  ", k["name"], k["ms"])
    for x in d.get("secondary", []): print("  sec", {k: x.get(k) for k in ("name","ms_per_step","value","launches_per_step")})
except Exception as e: print("no json", e)
PY
