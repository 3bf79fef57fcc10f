#!/bin/bash
# round 2, GPU session D: two-batch pipeline + lean loop for the single-buffer shape
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_scoring.py -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -n 3 gpurun_out/r2d_pytest.log
for m in DistMult; do
  timeout 300 python scripts/fullsort_probe.py --model $m --users 75776 --reps 6 --blocks 2 --path mma > gpurun_out/r2d_probe_$m.log 2>&1
  tail -n 2 gpurun_out/r2d_probe_$m.log
done
TAG=r2d_ bash scripts/gpu_exp_sweep.sh NMMA2 CHUNKED NOFILTER NOMMA 2>&1 | tee gpurun_out/r2d_exp.log
OUT=r2d_prof_sweep bash scripts/gpu_prof_sweep.sh
