#!/bin/bash
# One GPU-box pass: smoke, parity tests, bench lines of every config, ncu launch list, ncu --set full of the dominant kernel.
set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_b8192.json 2> gpurun_out/bench_b8192.err; echo "bench exit $?"
timeout 600 python bench.py --batch 100 --steps 500 --warmup 50 > gpurun_out/bench_b100.json 2> gpurun_out/bench_b100.err; echo "bench100 exit $?"
timeout 600 python bench.py --config conv --steps 50 --warmup 5 > gpurun_out/bench_conv.json 2> gpurun_out/bench_conv.err; echo "conv exit $?"
timeout 600 python bench.py --config scaled --steps 50 --warmup 5 > gpurun_out/bench_scaled.json 2> gpurun_out/bench_scaled.err; echo "scaled exit $?"
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:group -s 8 -c 4 -f -o gpurun_out/prof_group \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo done
