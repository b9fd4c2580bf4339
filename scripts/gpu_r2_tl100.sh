#!/bin/bash
mkdir -p gpurun_out
export VAEASSOC_LIB=$PWD/vae_assoc_b200/libvaeassoc_tl.so
VAEASSOC_TC_TIMELINE=1 VAEASSOC_TC_TIMELINE_ALL=1 timeout 300 python bench.py --batch 100 --steps 3 --warmup 3 --no-cpu-baseline --no-parity --no-secondary > gpurun_out/r2tl_100.json 2> gpurun_out/r2tl_100.txt
echo "exit $?"; grep -c "task" gpurun_out/r2tl_100.txt
