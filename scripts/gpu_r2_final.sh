#!/bin/bash
# round 2, final pass on one B200: full GPU suite, smoke, the default bench line (with secondary configs and the CPU
# baseline), the reference arm, an ncu launch list of a short bench run and an ncu --set full capture of the step kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2final_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2final_smoke.log 2>&1; echo "smoke exit $?"; tail -4 gpurun_out/r2final_smoke.log
timeout 900 python bench.py > gpurun_out/r2final_bench.json 2> gpurun_out/r2final_bench.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2final_bench_reference.json 2> gpurun_out/r2final_bench_reference.err; echo "reference arm exit $?"; cat gpurun_out/r2final_bench_reference.json | cut -c1-400
CMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-parity --no-secondary"
timeout 300 $CMD > gpurun_out/r2final_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2final_launches.csv $CMD > gpurun_out/r2final_ncu_launches.log 2>&1
echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_group_kernel -s 4 -c 1 -f -o gpurun_out/r2_step_kernel $CMD --quick > gpurun_out/r2final_ncu_full.log 2>&1
echo "ncu full exit $?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2final_bench.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step","gpu_launches")}, "e2e", d["e2e"]["value"], "dd", d["e2e_device_dataset"]["value"])
print("roofline", d["roofline"]["kernel"], d["roofline"]["frac"], "parity", d.get("parity"))
for k in d["kernels"][:10]: print("  ", k["name"], k["ms"])
for x in d.get("secondary", []): print("  sec", {k: x.get(k) for k in ("name","ms_per_step","value","launches_per_step")})
print("cpu", d.get("cpu_baseline"))
PY
