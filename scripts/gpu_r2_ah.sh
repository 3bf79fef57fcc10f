#!/bin/bash
# round 2, GPU session AH (1 GPU): the other learners + the whole train / distributed-free suite
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -k other_learners > gpurun_out/r2ah_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2ah_pytest.log
grep -E "^E  |passed|failed|FAILED|rc=" gpurun_out/r2ah_pytest.log | head -n 30 | cut -c1-300
