#!/bin/bash
# round 2, GPU session Q (1 GPU): TorusE through every test that runs over the model list + the new pipeline test
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2q_pytest.log
tail -n 30 gpurun_out/r2q_pytest.log | cut -c1-300
