#!/bin/bash
# round 2, GPU session C: lean chunk loop (no bitmap in the sweep, sign-bit group masks), shapes a / f for DistMult
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_scoring.py -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -n 4 gpurun_out/r2c_pytest.log
for m in DistMult ComplEx TransE; do
  timeout 300 python scripts/fullsort_probe.py --model $m --users 75776 --reps 6 --blocks 2 --path mma > gpurun_out/r2c_probe_$m.log 2>&1
  tail -n 2 gpurun_out/r2c_probe_$m.log
done
TAG=r2c_ bash scripts/gpu_exp_sweep.sh NMMA2 NOFILTER NOMMA 2>&1 | tee gpurun_out/r2c_exp.log
KGE_MMA_CFG=f TAG=r2c_f_ bash scripts/gpu_exp_sweep.sh 2>&1 | tee -a gpurun_out/r2c_exp.log
TAG=r2c_cx_ PROBE_ARGS="--model ComplEx" bash scripts/gpu_exp_sweep.sh 2>&1 | tee -a gpurun_out/r2c_exp.log
TAG=r2c_te_ PROBE_ARGS="--model TransE --d 100" bash scripts/gpu_exp_sweep.sh 2>&1 | tee -a gpurun_out/r2c_exp.log
OUT=r2c_prof_sweep bash scripts/gpu_prof_sweep.sh
