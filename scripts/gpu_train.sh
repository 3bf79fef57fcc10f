#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q --timeout 300 -x > gpurun_out/pytest_train.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_train.log); tail -4 gpurun_out/pytest_train.log
bash scripts/gpu_train_variants.sh "$@"
