#!/bin/bash
# quick pass: GEMM + parity tests, timing at 8192 / 1024 / 100 pairs, then the B = 100 timeline (timeline build)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/q_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/q_pytest.log
for B in 8192 1024 100; do echo -n "B=$B "; timeout 300 python bench.py --batch $B --steps 200 --warmup 20 --no-cpu-baseline --no-parity --no-secondary --quick 2>/dev/null | tail -1; done
bash scripts/gpu_r2_tl100.sh
