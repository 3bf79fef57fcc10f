#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q --timeout 120 -k "prefetcher" -x > gpurun_out/pytest_pf.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_pf.log); tail -4 gpurun_out/pytest_pf.log
python scripts/e2e_probe.py | tail -13
python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("bench ms/step", d["ms_per_step"], "e2e", d["e2e"])'
