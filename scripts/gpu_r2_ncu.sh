#!/bin/bash
# round 2: ncu --set full (with source-level stall samples) of ONE launch of the one-launch step kernel at 8192 pairs
mkdir -p gpurun_out
CMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-parity --no-secondary --quick"
timeout 300 $CMD > gpurun_out/r2n_plain.log 2>&1 && tail -c 300 gpurun_out/r2n_plain.log &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_group_kernel -s 4 -c 1 -f -o gpurun_out/r2_step_kernel $CMD > gpurun_out/r2n_ncu.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/r2n_ncu.log
