#!/bin/bash
# quick GPU pass: GEMM + parity tests, bench line, per-task timeline
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_quick.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_quick.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench exit $?"
VAEASSOC_TC_TIMELINE=1 VAEASSOC_TC_TIMELINE_ALL=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/timeline.txt; echo "tl exit $?"
