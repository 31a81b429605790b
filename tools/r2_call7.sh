#!/bin/bash
# round-2 call 7 (1 GPU): variant probes -> pick the x-row variant -> smoke, full GPU test suite, bench + ncu evidence (profiles/r02a_*)
mkdir -p gpurun_out
run() { echo "=== $1" >> gpurun_out/c7_probe.log; shift; env "$@" timeout 600 python tools/perf_probe.py --n 512 512 100 --fast 1 --reps 5 >> gpurun_out/c7_probe.log 2>&1; }
run "base" X=1
run "xpf (prefetch of p_old / M^-1 rows, 7 warps per SM)" NF_LIB=tools/_variants/xpf.so
run "ycol3, 2 CTAs per SM, carveout 66" NF_YCOL=3 NF_YCOL3_CTAS=2 NF_YCOL3_CARVE=66
run "ycol3, 3 CTAs per SM, carveout 100" NF_YCOL=3 NF_YCOL3_CARVE=100
grep -v "^problem built\|^upload\|sweep_\|cg_update\|cg_pupdate\|separate\|path \|slab_" gpurun_out/c7_probe.log
PICK=$(python - <<'PY'
import re
t=open('gpurun_out/c7_probe.log').read().split('=== ')[1:]
v={}
for b in t:
    m=re.search(r'cg_iteration\s+([0-9.]+) ms',b)
    if m: v[b.split('\n')[0]]=float(m.group(1))
base=[x for k,x in v.items() if k.startswith('base')][0]
xpf=[x for k,x in v.items() if k.startswith('xpf')][0]
print('xpf' if xpf < 0.98*base else 'base')
PY
)
echo "picked x-row variant: $PICK" | tee gpurun_out/c7_pick.log
if [ "$PICK" = "xpf" ]; then export NF_LIB=tools/_variants/xpf.so; fi
( time timeout 600 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/c7_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/c7_smoke.log | cut -c1-300
( time timeout 1800 python -m pytest tests -m gpu -x -q ) > gpurun_out/c7_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/c7_pytest.log
bash tools/profile_round.sh r02a --steps 5 --warmup 2
