#!/bin/bash
# usage: tools/sweep_mesh.sh "<mesh:states ...>" <batch>   (GPU box) - config-5 style sweep
for ms in $1; do
  m=${ms%%:*}; s=${ms##*:}
  echo "== mesh $m states $s batch $2"
  timeout 300 python tools/gpu_probe.py $m $s $2 3 2>&1 | tail -3
done
