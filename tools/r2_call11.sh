#!/bin/bash
# round-2 call 11 (gpurun --gpus 2): NCCL point-to-point settings for the neighbour exchange of the z-slab ranks
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
run() { tag=$1; shift; ( env "$@" timeout 400 $TR bench.py --gpus 2 --steps 2 --warmup 1 --no-converged --no-parity --mesh 512 512 100 ) > gpurun_out/c11_$tag.log 2>&1; echo "$tag rc=$?"; }
run default X=1
run p2pch NCCL_MIN_P2P_NCHANNELS=16 NCCL_MAX_P2P_NCHANNELS=32
run cememcpy NCCL_P2P_USE_CUDA_MEMCPY=1
run noovl NF_SLAB_OVERLAP=0
python - <<'PY'
import json
for t in ('default','p2pch','cememcpy','noovl'):
    for l in open(f'gpurun_out/c11_{t}.log'):
        if l.startswith('{'):
            d=json.loads(l); k=d['roofline']['kernels_ms']
            print(t,'value',round(d['value'],2),'cg_iteration',round(k['cg_iteration'],3),'xrow',round(k['xrow'],3),'ycol',round(k['ycol'],3),'zfwd+exch',round(k['zfwd'],3),'zback-phase',round(k['zback_update'],3),'keff',d['keff_after_K'])
PY
tail -2 gpurun_out/c11_cememcpy.log | cut -c1-300
