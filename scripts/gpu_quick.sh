#!/bin/bash
# quick GPU check: smoke, GPU tests, headline bench
mkdir -p gpurun_out
(timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log); tail -2 gpurun_out/smoke.log
(timeout 1200 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log); tail -15 gpurun_out/pytest.log
(timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err); tail -3 gpurun_out/bench.err; cat gpurun_out/bench.log
