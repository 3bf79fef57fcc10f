#!/bin/bash
python -m pytest tests/test_gpu_train.py tests/test_gpu_loader.py -m gpu -q -x 2>&1 | tail -2
python scripts/steps_probe.py
python scripts/steps_probe.py cfg3_rotate_yelp 2>&1 | tail -3
bash scripts/gpu_train_variants.sh 2>&1 | tail -3
