#!/bin/bash
# round 2, GPU session R (1 GPU): popularity / dynamic sampling through hopwise's loader; sweep shapes a / f / c
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sampler_pipeline.py -m gpu -q > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2r_pytest.log
grep -E "^E  |passed|failed|differ" gpurun_out/r2r_pytest.log | head -n 20
for cfg in a f c; do
  for m in DistMult ComplEx; do
    echo "cfg=$cfg $m: $(KGE_MMA_CFG=$cfg timeout 120 python scripts/fullsort_probe.py --users 75776 --reps 5 --path mma --model $m 2>&1 | tail -n 1)"
  done
done
