#!/bin/bash
# Development tool: device-time sweep of the fused CG iteration's tuning knobs (run on the GPU box through gpurun).
# usage: tools/fused_sweep.sh NX NY NZ  "ENV1=.. ENV2=.." "ENV..." ...
nx=$1; ny=$2; nz=$3; shift 3
for cfg in "$@"; do
  echo "=== $nx x $ny x $nz :: $cfg"
  env $cfg timeout 600 python tools/perf_probe.py --n $nx $ny $nz --fast 1 --reps 5 2>&1 | grep -v "^problem built\|^upload"
done
