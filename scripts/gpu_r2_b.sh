#!/bin/bash
# round 2, GPU session B: the "span" epilogue (warps alternate between the two accumulators of a CTA)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_scoring.py -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -n 4 gpurun_out/r2b_pytest.log
for m in DistMult ComplEx TransE; do
  timeout 300 python scripts/fullsort_probe.py --model $m --users 75776 --reps 6 --blocks 2 --path mma > gpurun_out/r2b_probe_$m.log 2>&1
  tail -n 2 gpurun_out/r2b_probe_$m.log
done
TAG=r2b_ bash scripts/gpu_exp_sweep.sh NMMA2 MERGED NOFILTER NOMMA 2>&1 | tee gpurun_out/r2b_exp.log
TAG=r2b_cx_ PROBE_ARGS="--model ComplEx" bash scripts/gpu_exp_sweep.sh 2>&1 | tee -a gpurun_out/r2b_exp.log
TAG=r2b_te_ PROBE_ARGS="--model TransE --d 100" bash scripts/gpu_exp_sweep.sh 2>&1 | tee -a gpurun_out/r2b_exp.log
OUT=r2b_prof_sweep bash scripts/gpu_prof_sweep.sh
