#!/bin/bash
# Development tool: one GPU-box session = parity tests + per-path kernel timings + the bench line.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
for m in 0 1 2; do
  echo "=== NF_FUSED=$m" >> gpurun_out/probe.log
  NF_FUSED=$m timeout 600 python tools/perf_probe.py --n 256 256 200 --fast 1 --reps 5 >> gpurun_out/probe.log 2>&1
done
cat gpurun_out/probe.log
( time timeout 1500 python bench.py ) > gpurun_out/bench_full.log 2>&1
tail -3 gpurun_out/bench_full.log
