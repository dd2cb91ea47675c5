#!/bin/bash
# usage: tools/sweep_env.sh <mesh> <states> "<batches>" "<chunk_xts>"   (tuning helper, GPU box)
for b in $3; do for c in $4; do
  echo "== batch $b chunk_xt $c"
  CPB_CHUNK_XT=$c timeout 300 python tools/gpu_probe.py $1 $2 $b 3 2>&1 | tail -2
done; done
