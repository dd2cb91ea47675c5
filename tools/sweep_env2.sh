#!/bin/bash
# usage: tools/sweep_env2.sh <mesh> <states> "<batches>" "<ENV=val ...>" ...   (tuning helper, GPU box)
m=$1; s=$2; bs=$3; shift 3
for b in $bs; do for e in "$@"; do
  echo "== batch $b env $e"
  env $e timeout 300 python tools/gpu_probe.py $m $s $b 3 2>&1 | tail -3
done; done
