#!/bin/bash
# quick pass: parity tests, then the HBM-resident timing loop at 8192 / 100 pairs under a few scheduling switches
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused_schedules or reference_config or ragged or 8192" > gpurun_out/q_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/q_pytest.log
run() { for B in 8192 100; do echo -n "$1 B=$B "; env $1 timeout 300 python bench.py --batch $B --steps 200 --warmup 20 --no-cpu-baseline --no-parity --no-secondary --quick 2>/dev/null | tail -1; done; }
run "X=1"
run "VAEASSOC_ORDER_DECLARED=1"
run "VAEASSOC_NO_HALF=1"
