#!/bin/bash
# quick pass: a few parity tests, then the HBM-resident timing loop at 8192 and 100 pairs
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused_schedules or reference_config or ragged" > gpurun_out/q_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/q_pytest.log
for B in 8192 100; do
  timeout 300 python bench.py --batch $B --steps 200 --warmup 20 --no-cpu-baseline --no-parity --no-secondary --quick 2>/dev/null | tail -1
done
# throughput bound of the tile engine: the four-segment schedule WITHOUT dependency waits (wrong results, timing only)
VAEASSOC_DEBUG_NODEPS=1 timeout 300 python bench.py --batch 8192 --steps 100 --warmup 20 --no-cpu-baseline --no-parity --no-secondary 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('NODEPS', d['ms_per_step'], [(k['name'], k['ms']) for k in d['kernels'][:12]])"
