#!/bin/bash
# round-2 development call 6 (gpurun --gpus 2): neighbour-exchange mode of thick z-slabs
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_gpu_slab.py -x -q ) > gpurun_out/c6_slab.log 2>&1; echo "slab tests rc=$?"; tail -40 gpurun_out/c6_slab.log | cut -c1-220
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
( time timeout 900 $TR bench.py --gpus 2 --steps 3 --warmup 1 --no-converged --no-parity ) > gpurun_out/c6_bench2.log 2>&1; echo "bench2 rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/c6_bench2.log',):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l)
            print(f,'value',round(d['value'],2),'e2e',round(d['e2e']['value'],2),'frac',round(d['roofline']['frac'],3),'in_run',round(d['roofline']['in_run']['frac'],3))
            print('  kernels',{k:round(v,3) for k,v in d['roofline']['kernels_ms'].items()})
PY
tail -3 gpurun_out/c6_bench2.log | cut -c1-300
