#!/bin/bash
# round-2 call 8 (gpurun --gpus 2): z-slab tests after the s0 cut, config-4 golden test, 2-GPU bench line
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_slab.py "tests/test_gpu_keff.py::test_config4_koeberg_34x34_golden" -x -q ) > gpurun_out/c8_tests.log 2>&1; echo "tests rc=$?"; tail -25 gpurun_out/c8_tests.log | cut -c1-220
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
( time timeout 900 $TR bench.py --gpus 2 --steps 3 --warmup 1 --no-converged ) > gpurun_out/r02a_bench2.log 2>&1; echo "bench2 rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r02a_bench2.log'):
    if l.startswith('{'):
        d=json.loads(l)
        print('N=2 value',round(d['value'],2),'e2e',round(d['e2e']['value'],2),'frac',round(d['roofline']['frac'],3),'in_run',round(d['roofline']['in_run']['frac'],3))
        print('  kernels',{k:round(v,3) for k,v in d['roofline']['kernels_ms'].items()})
        print('  parity',d.get('parity_vs_n1'))
PY
