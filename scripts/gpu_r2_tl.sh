#!/bin/bash
# round 2: per-task timeline (%globaltimer stamps; timeline build of the library: python -m vae_assoc_b200.build --timeline)
# of the one-launch step at 8192 and 100 pairs -> gpurun_out/r2tl_<B>.txt, digested by scripts/timeline_digest.py
mkdir -p gpurun_out
export VAEASSOC_LIB=$PWD/vae_assoc_b200/libvaeassoc_tl.so
for B in ${@:-8192 100}; do
  VAEASSOC_TC_TIMELINE=1 VAEASSOC_TC_TIMELINE_ALL=1 timeout 300 python bench.py --batch $B --steps 3 --warmup 3 --no-cpu-baseline --no-parity --no-secondary > gpurun_out/r2tl_$B.json 2> gpurun_out/r2tl_$B.txt
  echo "B=$B exit $?"; grep -c "task" gpurun_out/r2tl_$B.txt
done
