#!/bin/bash
# quick pass: GEMM unit tests + parity tests, then the HBM-resident timing loop at 8192 / 1024 / 100 / 64 pairs, with and
# without the switch(es) given as arguments (e.g. VAEASSOC_NO_SMALL_ROWS=1)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_parity.py tests/test_gpu_callers.py -m gpu -q -x > gpurun_out/q_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/q_pytest.log
run() { for B in 8192 1024 200 100 64; do echo -n "$1 B=$B "; env $1 timeout 300 python bench.py --batch $B --steps 200 --warmup 20 --no-cpu-baseline --no-parity --no-secondary --quick 2>/dev/null | tail -1; done; }
run "X=1"
if [ -n "$1" ]; then run "$1"; fi
