#!/bin/bash
# round 2: where the tile epilogue's time goes -- four-segment schedule without dependency waits (pure throughput of the
# tile engine; wrong results, timing only) with parts of the NN / NT epilogue switched off
mkdir -p gpurun_out
# (the SKIP / TMA-store switches live in the debug build: python -m vae_assoc_b200.build --epi-debug)
export VAEASSOC_LIB=$PWD/vae_assoc_b200/libvaeassoc_dbg.so
run() {
  env VAEASSOC_DEBUG_NODEPS=1 $1 timeout 300 python bench.py --batch 8192 --steps 50 --warmup 10 --no-cpu-baseline --no-parity --no-secondary 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
k = {x['name']: x['ms'] for x in d['kernels']}
print('$1', 'step %.4f' % d['ms_per_step'], ' '.join('%s %.1f' % (n, 1e3 * k[n]) for n in ('seg_fwd_enc', 'seg_fwd_dec', 'seg_bwd_dec', 'seg_bwd_enc') if n in k))"
}
run "X=1"
run "VAEASSOC_DEBUG_SKIP_MATH=1"
run "VAEASSOC_DEBUG_SKIP_STORE=1"
run "VAEASSOC_DEBUG_SKIP_MATH=1 VAEASSOC_DEBUG_SKIP_STORE=1"
run "VAEASSOC_EPI_TMA_STORE=1"
run "VAEASSOC_NO_MASK=1"
