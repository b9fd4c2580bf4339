#!/bin/bash
# compute-sanitizer over the hand-rolled release/acquire + async-proxy protocol of the tile kernel (gemm_group.cu) and the
# rest of the step: ONE tool per call (scripts/sanitize.sh memcheck|racecheck|synccheck) on two train steps of the
# one-launch tf32 schedule at B = 1024 (four row blocks: 128-wide tiles, half-tile hand-over, elementwise tasks, signal
# warp).  Log -> gpurun_out/r2_sanitizer_<tool>.log ; summaries are copied to profiles/ by hand.
TOOL=${1:-memcheck}
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 300 python scripts/sanitize_step.py 1024 > gpurun_out/r2_sanitizer_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r2_sanitizer_plain.log; exit 1; }
tail -1 gpurun_out/r2_sanitizer_plain.log
timeout 1200 $CS --tool $TOOL --print-limit 20 python scripts/sanitize_step.py 1024 > gpurun_out/r2_sanitizer_$TOOL.log 2>&1
echo "exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|costs|Error|error" gpurun_out/r2_sanitizer_$TOOL.log | head -12
