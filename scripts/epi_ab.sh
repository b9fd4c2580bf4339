#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 400 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_quick.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/pytest_quick.log
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/bench_epi_strip.json 2>/dev/null; echo "strip $?"
VAEASSOC_EPI_BIAS_SHFL=1 timeout 200 python bench.py --no-cpu-baseline > gpurun_out/bench_epi_shfl.json 2>/dev/null; echo "shfl $?"
VAEASSOC_EPI_TMA_STORE=1 timeout 200 python bench.py --no-cpu-baseline > gpurun_out/bench_epi_tma.json 2>/dev/null; echo "tma $?"
VAEASSOC_EPI_TMA_STORE=1 timeout 300 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_tma.log 2>&1; echo "pytest tma exit $?"; tail -2 gpurun_out/pytest_tma.log
