#!/bin/bash
# round-1 GPU session B: full test run, bench, tcgen05 full-sort profile
mkdir -p gpurun_out
(timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log); tail -2 gpurun_out/smoke.log
(timeout 1200 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log); tail -15 gpurun_out/pytest.log
(timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err); tail -3 gpurun_out/bench.err; cat gpurun_out/bench.log
F="python scripts/fullsort_probe.py --users 37888 --reps 2 --path mma"
$F > gpurun_out/plain_fs_mma.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"fullsort_mma|rescore_topk" -s 2 -c 2 -o gpurun_out/prof_fullsort_mma $F > gpurun_out/ncu_fs_mma.log 2>&1
cat gpurun_out/plain_fs_mma.log | tail -3
python scripts/fullsort_probe.py --users 37888 --reps 2 --path mma --model ComplEx | tail -2
python scripts/fullsort_probe.py --users 37888 --reps 2 --path mma --model TransE --d 100 | tail -2
