#!/bin/bash
# compute-sanitizer over the hand-rolled release/acquire + async-proxy protocol of the tile kernel (gemm_group.cu) and the
# rest of the step: memcheck + synccheck + racecheck on one fused tf32 step at B = 512, memcheck on the GEMM unit tests.
# Logs (tails) -> gpurun_out/r2_sanitizer_*.log ; summaries are copied to profiles/ by hand.
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck synccheck racecheck; do
  echo "== $tool: fused step B=512"
  timeout 900 $CS --tool $tool --print-limit 20 python scripts/sanitize_step.py 512 > gpurun_out/r2_sanitizer_step_$tool.log 2>&1
  echo "exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|costs" gpurun_out/r2_sanitizer_step_$tool.log | tail -3
done
echo "== memcheck: tests/test_gpu_gemm.py"
timeout 1500 $CS --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_gemm.py -m gpu -q -x > gpurun_out/r2_sanitizer_gemm_memcheck.log 2>&1
echo "exit $?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2_sanitizer_gemm_memcheck.log | tail -3
