#!/bin/bash
# round 2, pass A: full GPU test suite (order-proof), smoke, headline bench, tile-width experiment
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -s > gpurun_out/r2a_pytest.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
grep -E "\[masks\]|\[tf32 order|\[latents_1k\]|^ +[0-9]+ +[0-9.e+-]+ " gpurun_out/r2a_pytest.log | head -40
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1; echo "smoke exit $?"; tail -4 gpurun_out/r2a_smoke.log
timeout 900 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench exit $?"; tail -3 gpurun_out/r2a_bench.err
for bn in 128 192; do
  VAEASSOC_MAX_BN=$bn timeout 300 python bench.py --quick --steps 100 > gpurun_out/r2a_bn$bn.json 2> gpurun_out/r2a_bn$bn.err; echo "bn$bn exit $?"; cat gpurun_out/r2a_bn$bn.json
done
timeout 300 python bench.py --quick --steps 100 > gpurun_out/r2a_bn256.json 2>/dev/null; cat gpurun_out/r2a_bn256.json
