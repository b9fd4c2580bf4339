#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "two_launch or ragged" > gpurun_out/r2c_pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/r2c_pytest.log
for v in none VAEASSOC_DEBUG_SKIP_MATH VAEASSOC_DEBUG_SKIP_STORE both; do
  if [ $v = none ]; then E=""; elif [ $v = both ]; then E="VAEASSOC_DEBUG_SKIP_MATH=1 VAEASSOC_DEBUG_SKIP_STORE=1"; else E="$v=1"; fi
  echo "== $v"; env $E timeout 300 python bench.py --quick --steps 200 2>/dev/null
  env $E timeout 300 python bench.py --quick --steps 400 --batch 100 2>/dev/null
done
