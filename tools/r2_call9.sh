#!/bin/bash
# round-2 call 9 (gpurun --gpus 8): final 8-GPU and 4-GPU bench lines
mkdir -p gpurun_out
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523"
( time timeout 600 $TR8 bench.py --gpus 8 --steps 5 --warmup 2 ) > gpurun_out/r02a_bench8.log 2>&1; echo "bench8 rc=$?"
( time timeout 400 $TR4 bench.py --gpus 4 --steps 3 --warmup 1 --no-converged --no-parity ) > gpurun_out/r02a_bench4.log 2>&1; echo "bench4 rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/r02a_bench8.log','gpurun_out/r02a_bench4.log'):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l)
            print(f,'value',round(d['value'],2),'e2e',round(d['e2e']['value'],2),'frac',round(d['roofline']['frac'],3),'in_run',round(d['roofline']['in_run']['frac'],3),'ms/step',round(d['ms_per_step'],1))
            print('  kernels',{k:round(v,3) for k,v in d['roofline']['kernels_ms'].items()})
            print('  parity',d.get('parity_vs_n1'))
            print('  ttk',json.dumps(d.get('time_to_keff'))[:900])
PY
