#!/bin/bash
# fp32-exact SIMT path: GEMM unit tests (use_tc = 0 and 1), the fp32 parity / reproducibility tests, then its step time
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_parity.py tests/test_gpu_conv.py -m gpu -q -x > gpurun_out/fp32_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/fp32_pytest.log
for e in X=1 VAEASSOC_SIMT_SMALL_TILES=1; do echo -n "$e "; env $e timeout 300 python bench.py --precision fp32 --steps 12 --warmup 5 --no-cpu-baseline --no-parity --no-secondary --quick 2>/dev/null | tail -1; done
