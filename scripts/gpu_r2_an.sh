#!/bin/bash
mkdir -p gpurun_out
timeout 100 python -m pytest tests -m gpu -x -q > gpurun_out/r2an_pytest.log 2>&1; echo "pytest rc=$?"
tail -n 3 gpurun_out/r2an_pytest.log | cut -c1-200
