#!/bin/bash
# usage: tools/sweep_libs.sh "<lib suffixes>" <mesh> <states> <batch> [chunk_xt]   (tuning helper, GPU box)
for v in $1; do
  lib=$PWD/cpmd_b200/libcpb200$v.so
  echo "== lib$v mesh $2 states $3 batch $4 chunk ${5:-24}"
  CPB200_LIB=$lib CPB_CHUNK_XT=${5:-24} timeout 300 python tools/gpu_probe.py $2 $3 $4 2>&1 | tail -2
done
