#!/bin/bash
# round 2, GPU session AC (1 GPU): whole GPU suite with TransD
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2ac_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2ac_pytest.log
grep -E "^E  |passed|failed|FAILED|rc=" gpurun_out/r2ac_pytest.log | head -n 40 | cut -c1-300
