#!/bin/bash
# round 2, pass E: full GPU suite after the guard / inference-graph / two-launch changes, then the full default bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2e_pytest.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/r2e_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2e_smoke.log 2>&1; echo "smoke exit $?"; tail -4 gpurun_out/r2e_smoke.log
timeout 900 python bench.py > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench exit $?"; tail -3 gpurun_out/r2e_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2e_bench.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step","gpu_launches")}, "e2e", d["e2e"]["value"], "dd", d["e2e_device_dataset"]["value"])
print("roofline", d["roofline"]["kernel"], d["roofline"]["frac"], "parity", d.get("parity"))
for k in d["kernels"][:10]: print("  ", k["name"], k["ms"])
for x in d.get("secondary", []): print("  sec", x)
PY
