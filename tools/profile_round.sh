#!/bin/bash
# One GPU-box session that produces everything profiles/ needs for a round: the bench line, the ncu launch list of the
# same command, and one ncu --set full capture of each kernel of the CG iteration at the bench workload.
# usage: tools/profile_round.sh TAG [bench args]
tag=$1; shift
mkdir -p gpurun_out
( time timeout 1500 python bench.py "$@" ) > gpurun_out/${tag}_bench.log 2>&1 || exit 1
tail -4 gpurun_out/${tag}_bench.log | cut -c1-600
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-converged "$@" > gpurun_out/${tag}_ncu_launches.log 2>&1
# the --set full captures replay each kernel ~45 times and save / restore the device memory it touches in between: they run
# on a quarter of the planes (same plane size, same kernels; DRAM bytes per DOF do not depend on nz) to keep that affordable
FULL_MESH=${FULL_MESH:-"512 512 100"}
for k in k_xrow k_ycol k_zfwd k_zback_update; do
  out=gpurun_out/${tag}_full_${k}
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$k" --launch-skip 20 --launch-count 1 -f -o $out \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-converged --mesh $FULL_MESH > $out.log 2>&1
  ncu -i $out.ncu-rep --page raw --csv > $out.raw.csv 2>/dev/null
  rm -f $out.ncu-rep
done
ls -la gpurun_out | tail -12
