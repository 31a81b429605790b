#!/bin/bash
# round-2 development call 2: vectorised z kernels, outer accelerators, first full-size bench line
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_outputs.py tests/test_gpu_keff.py -x -q ) > gpurun_out/c2_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/c2_tests.log
for v in base zv1 zv2 zv3; do
  lib=tools/_variants/$v.so; [ $v = base ] && lib=neutfem_b200/lib/libneutfem_b200.so
  echo "=== variant $v" >> gpurun_out/c2_probe.log
  NF_LIB=$lib timeout 600 python tools/perf_probe.py --n 512 512 100 --fast 1 --reps 5 >> gpurun_out/c2_probe.log 2>&1
done
echo "=== base, parity mode" >> gpurun_out/c2_probe.log
timeout 600 python tools/perf_probe.py --n 512 512 100 --fast 0 --reps 5 >> gpurun_out/c2_probe.log 2>&1
grep -v "^problem built\|^upload\|sweep_\|cg_update\|cg_pupdate\|separate\|path " gpurun_out/c2_probe.log
for cfg in "--accel chebyshev" "--accel anderson" "--accel chebyshev --eta 0.1" "--accel chebyshev --eta 0.03" "--accel anderson --eta 0.1" "--accel none"; do
  echo "=== 256x256x200 converged: $cfg" >> gpurun_out/c2_outer.log
  timeout 600 python tools/perf_probe.py --n 256 256 200 --fast 1 --no-kernels --outer 300 $cfg >> gpurun_out/c2_outer.log 2>&1
done
grep "===\|solve_keff" gpurun_out/c2_outer.log
( time timeout 1200 python bench.py --steps 3 --warmup 1 --no-converged ) > gpurun_out/c2_bench.log 2>&1; echo "bench rc=$?"; tail -2 gpurun_out/c2_bench.log | cut -c1-1500
