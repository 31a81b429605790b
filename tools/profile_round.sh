#!/bin/bash
# One GPU-box session that produces everything profiles/ needs for a round: the bench line, the ncu launch list of the
# same command, and one ncu --set full capture of each kernel of the CG iteration at the plane size of the bench workload.
# usage: tools/profile_round.sh TAG [bench args]
tag=$1; shift
mkdir -p gpurun_out
( time timeout 1500 python bench.py "$@" ) > gpurun_out/${tag}_bench.log 2>&1 || exit 1
tail -4 gpurun_out/${tag}_bench.log | cut -c1-600
LCMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-converged"
timeout 900 $LCMD > gpurun_out/${tag}_plain_launch.log 2>&1 &&
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    $LCMD > gpurun_out/${tag}_ncu_launches.log 2>&1
# the --set full captures replay each kernel ~45 times and save / restore the device memory it touches in between: they run
# on a quarter of the planes (same plane size, same kernel variants; DRAM bytes per DOF do not depend on nz)
FULL_MESH=${FULL_MESH:-"512 512 100"}
PCMD="python tools/perf_probe.py --n $FULL_MESH --fast 1 --reps 2"
timeout 600 $PCMD > gpurun_out/${tag}_plain_probe.log 2>&1 || exit 1
for k in ${KERNELS:-k_xrow k_ycol k_zfwd2 k_zback2}; do
  out=gpurun_out/${tag}_full_${k}
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$k" --launch-skip 3 --launch-count 1 -f -o $out \
      $PCMD > $out.log 2>&1
  ncu -i $out.ncu-rep --page raw --csv > $out.raw.csv 2>/dev/null
  ncu -i $out.ncu-rep --page source --csv > $out.source.csv 2>/dev/null
  rm -f $out.ncu-rep
done
ls -la gpurun_out | tail -14
