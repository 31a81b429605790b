#!/bin/bash
# round-2 development call 3: three-phase y columns, lagged polling, full default bench
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_fused.py tests/test_gpu_accel.py tests/test_gpu_outputs.py tests/test_gpu_dropin.py -x -q ) > gpurun_out/c3_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/c3_tests.log
run() { echo "=== $1" >> gpurun_out/c3_probe.log; shift; env "$@" timeout 600 python tools/perf_probe.py --n 512 512 100 --fast 1 --reps 5 >> gpurun_out/c3_probe.log 2>&1; }
run "base (ycol3)" X=1
run "NF_YCOL=0 (old y columns)" NF_YCOL=0
run "yb8" NF_LIB=tools/_variants/yb8.so
run "zf5" NF_LIB=tools/_variants/zf5.so
grep -v "^problem built\|^upload\|sweep_\|cg_update\|cg_pupdate\|separate\|path " gpurun_out/c3_probe.log
( time timeout 1500 python bench.py --steps 5 --warmup 2 ) > gpurun_out/c3_bench.log 2>&1; echo "bench rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/c3_bench.log'):
    if l.startswith('{'):
        d=json.loads(l)
        print('value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],'in_run',d['roofline']['in_run'])
        print('ttk',json.dumps(d['time_to_keff'])[:1500])
        print('kernels',d['roofline']['kernels_ms'])
PY
tail -4 gpurun_out/c3_bench.log | cut -c1-300
