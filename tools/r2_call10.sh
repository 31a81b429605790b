#!/bin/bash
# round-2 call 10 (1 GPU): vectorised x-row passes -- parity tests, A/B against the scalar passes, bench + ncu evidence (profiles/r02b_*)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_fullsize.py "tests/test_gpu_keff.py::test_config4_koeberg_34x34_golden" -x -q ) > gpurun_out/c10_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/c10_tests.log
run() { echo "=== $1" >> gpurun_out/c10_probe.log; shift; env "$@" timeout 600 python tools/perf_probe.py --n 512 512 100 --fast 1 --reps 5 >> gpurun_out/c10_probe.log 2>&1; }
run "base (x-row passes on pairs of cells)" X=1
run "xv0 (scalar passes)" NF_LIB=tools/_variants/xv0.so
echo "=== base, parity mode" >> gpurun_out/c10_probe.log; timeout 600 python tools/perf_probe.py --n 512 512 100 --fast 0 --reps 5 >> gpurun_out/c10_probe.log 2>&1
grep -v "^problem built\|^upload\|sweep_\|cg_update\|cg_pupdate\|separate\|path \|slab_" gpurun_out/c10_probe.log
PICK=$(python - <<'PY'
import re
t=open('gpurun_out/c10_probe.log').read().split('=== ')[1:]
v={}
for b in t:
    m=re.search(r'cg_iteration\s+([0-9.]+) ms',b)
    if m: v[b.split('\n')[0]]=float(m.group(1))
base=[x for k,x in v.items() if k.startswith('base (x-row')][0]
xv0=[x for k,x in v.items() if k.startswith('xv0')][0]
print('xv0' if xv0 < 0.99*base else 'base')
PY
)
echo "picked: $PICK" | tee gpurun_out/c10_pick.log
if [ "$PICK" = "xv0" ]; then export NF_LIB=tools/_variants/xv0.so; fi
bash tools/profile_round.sh r02b --steps 5 --warmup 2
