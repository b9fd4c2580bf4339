#!/bin/bash
# round 2, pass B: two-launch schedule (latent stages as tasks of the tile kernel): parity, then bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2b_pytest.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/r2b_pytest.log
timeout 300 python bench.py --quick --steps 200 > gpurun_out/r2b_quick.json 2> gpurun_out/r2b_quick.err; echo "quick exit $?"; cat gpurun_out/r2b_quick.json
VAEASSOC_NO_ELT=1 timeout 300 python bench.py --quick --steps 200 > gpurun_out/r2b_quick_noelt.json 2>/dev/null; cat gpurun_out/r2b_quick_noelt.json
timeout 300 python bench.py --quick --steps 400 --batch 100 > gpurun_out/r2b_quick_b100.json 2>/dev/null; cat gpurun_out/r2b_quick_b100.json
VAEASSOC_NO_ELT=1 timeout 300 python bench.py --quick --steps 400 --batch 100 > gpurun_out/r2b_quick_b100_noelt.json 2>/dev/null; cat gpurun_out/r2b_quick_b100_noelt.json
