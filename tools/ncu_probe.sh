#!/bin/bash
# Development tool: ncu --set full captures of the hot kernels at a probe mesh (one launch each), run on the GPU box.
# usage: tools/ncu_probe.sh TAG NX NY NZ "kernel_regex[:skip]" ...   (env NF_FUSED etc. is inherited)
tag=$1; nx=$2; ny=$3; nz=$4; shift 4
mkdir -p gpurun_out
for spec in "$@"; do
  k=${spec%%:*}; skip=1; [[ "$spec" == *:* ]] && skip=${spec##*:}
  out=gpurun_out/${tag}_${k//[^A-Za-z0-9_]/}
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$k" --launch-skip $skip --launch-count 1 \
      -f -o $out python tools/perf_probe.py --n $nx $ny $nz --fast 1 --reps 2 > $out.log 2>&1
  ncu -i $out.ncu-rep --page raw --csv > $out.raw.csv 2>/dev/null
  tail -2 $out.log
done
