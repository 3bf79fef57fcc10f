#!/bin/bash
# round 2, GPU session P (1 GPU): id prefetch in the forward kernel -- default (two-per-warp shape only) vs off vs all
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q -x 2>&1 | tail -n 2
bash scripts/gpu_train_variants.sh pf0 pf1 2>&1 | tee gpurun_out/r2p_variants.log
