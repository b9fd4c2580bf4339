#!/bin/bash
for ns in 0 200 400 600 900; do
  echo "== stagger $ns"; VAEASSOC_EPI_STAGGER_NS=$ns timeout 300 python bench.py --quick --steps 200 2>/dev/null
done
