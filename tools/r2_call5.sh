#!/bin/bash
# round-2 development call 5 (gpurun --gpus 8): 8-GPU bench line (with converged solves and N-vs-1 parity), overlap A/B, 4-rank slab test
mkdir -p gpurun_out
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523"
( time timeout 900 $TR8 bench.py --gpus 8 --steps 5 --warmup 2 ) > gpurun_out/c5_bench8.log 2>&1; echo "bench8 rc=$?"
( time NF_SLAB_OVERLAP=0 timeout 600 $TR8 bench.py --gpus 8 --steps 3 --warmup 1 --no-converged --no-parity ) > gpurun_out/c5_bench8_noovl.log 2>&1; echo "bench8 (no overlap) rc=$?"
( time timeout 600 $TR4 bench.py --gpus 4 --steps 3 --warmup 1 --no-converged ) > gpurun_out/c5_bench4.log 2>&1; echo "bench4 rc=$?"
( time timeout 900 python -m pytest tests/test_gpu_slab.py -x -q -k "4" ) > gpurun_out/c5_slab4.log 2>&1; echo "slab4 rc=$?"; tail -3 gpurun_out/c5_slab4.log
python - <<'PY'
import json
for f in ('gpurun_out/c5_bench8.log','gpurun_out/c5_bench8_noovl.log','gpurun_out/c5_bench4.log'):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l)
            print(f,'value',round(d['value'],2),'e2e',round(d['e2e']['value'],2),'frac',round(d['roofline']['frac'],3),'in_run',round(d['roofline']['in_run']['frac'],3),'ms/step',round(d['ms_per_step'],1))
            print('  kernels',{k:round(v,3) for k,v in d['roofline']['kernels_ms'].items()})
            print('  parity',d.get('parity_vs_n1'))
            print('  ttk',json.dumps(d.get('time_to_keff'))[:900])
PY
tail -4 gpurun_out/c5_bench8.log | cut -c1-300
