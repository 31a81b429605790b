#!/bin/bash
# round-2 development call: smoke + new-path parity tests + kernel timings of the variants
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/c1_smi.log 2>&1
( time timeout 600 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/c1_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/c1_smoke.log
( time timeout 900 python -m pytest tests/test_gpu_fused.py -x -q ) > gpurun_out/c1_fused.log 2>&1; echo "fused rc=$?"; tail -6 gpurun_out/c1_fused.log
for v in base xw2 zb3 zb6; do
  lib=tools/_variants/$v.so; [ $v = base ] && lib=neutfem_b200/lib/libneutfem_b200.so
  echo "=== variant $v" >> gpurun_out/c1_probe.log
  NF_LIB=$lib timeout 600 python tools/perf_probe.py --n 512 512 100 --fast 1 --reps 5 >> gpurun_out/c1_probe.log 2>&1
done
echo "=== base, parity mode" >> gpurun_out/c1_probe.log
timeout 600 python tools/perf_probe.py --n 512 512 100 --fast 0 --reps 5 >> gpurun_out/c1_probe.log 2>&1
echo "=== base, per-lane feeding" >> gpurun_out/c1_probe.log
NF_XROW_BULK=0 timeout 600 python tools/perf_probe.py --n 512 512 100 --fast 1 --reps 5 >> gpurun_out/c1_probe.log 2>&1
for cap in 6 7; do
  echo "=== base, NF_XROW_CTAS=$cap" >> gpurun_out/c1_probe.log
  NF_XROW_CTAS=$cap timeout 600 python tools/perf_probe.py --n 512 512 100 --fast 1 --reps 5 >> gpurun_out/c1_probe.log 2>&1
done
grep -v "^problem built\|^upload" gpurun_out/c1_probe.log
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/c1_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/c1_pytest.log
