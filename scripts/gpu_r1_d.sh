#!/bin/bash
# round-1 GPU session D: smoke, full GPU test run, both bench arms (with the CPU baseline leg)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.log 2>&1
nproc >> gpurun_out/smi.log
(timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log); tail -2 gpurun_out/smoke.log
(timeout 1200 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log); tail -15 gpurun_out/pytest.log
(timeout 900 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err); tail -3 gpurun_out/bench.err; cat gpurun_out/bench.log
(timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "ref rc=$?" >> gpurun_out/bench_ref.err); tail -3 gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.log
