#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "default pytest exit $?"; tail -1 gpurun_out/pytest_gpu.log
timeout 60 python bench.py --no-cpu-baseline > gpurun_out/bench_w8.json 2>/dev/null; echo "w8 bench $?"
export VAEASSOC_LIB=$PWD/vae_assoc_b200/libvaeassoc_w16.so
timeout 90 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_w16.log 2>&1; echo "w16 pytest exit $?"; tail -1 gpurun_out/pytest_w16.log
timeout 60 python bench.py --no-cpu-baseline > gpurun_out/bench_w16.json 2>/dev/null; echo "w16 bench $?"
timeout 60 python bench.py --batch 100 --steps 500 --warmup 50 --no-cpu-baseline > gpurun_out/bench_w16_b100.json 2>/dev/null; echo "w16 b100 $?"
