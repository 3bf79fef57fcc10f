#!/bin/bash
# round 2, GPU session A: parity of the reworked tensor-core sweep, timings per variant, one full ncu capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/r2a_smi.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -n 5 gpurun_out/r2a_pytest.log
for m in DistMult ComplEx TransE; do
  timeout 300 python scripts/fullsort_probe.py --model $m --users 75776 --reps 8 --blocks 2 --path mma > gpurun_out/r2a_probe_$m.log 2>&1
  tail -n 4 gpurun_out/r2a_probe_$m.log
done
bash scripts/gpu_exp_sweep.sh MERGED NOFILTER NOMMA NOLDF 2>&1 | tee gpurun_out/r2a_exp.log
PROBE_ARGS="--model ComplEx" bash scripts/gpu_exp_sweep.sh 2>&1 | sed 's/default/ComplEx/' | tee -a gpurun_out/r2a_exp.log
OUT=r2a_prof_sweep bash scripts/gpu_prof_sweep.sh
