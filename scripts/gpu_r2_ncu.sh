#!/bin/bash
# round 2: ncu --set full (with source-level stall samples) of ONE launch of the one-launch step kernel at 8192 pairs
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity --no-secondary --quick"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused_schedules or reference_config or ragged" > gpurun_out/r2n_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2n_pytest.log
timeout 300 $CMD > gpurun_out/r2n_plain.log 2>&1 && tail -c 600 gpurun_out/r2n_plain.log &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_group_kernel -s 5 -c 1 -f -o gpurun_out/r2_step_kernel $CMD > gpurun_out/r2n_ncu.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/r2n_ncu.log
