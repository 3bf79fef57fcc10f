#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_scoring.py -m gpu -q --timeout 120 -k "mma" > gpurun_out/pytest_mma.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_mma.log); tail -8 gpurun_out/pytest_mma.log
python scripts/fullsort_probe.py --users 37888 --reps 4 --path mma | tail -3
python scripts/fullsort_probe.py --users 37888 --reps 6 --path mma --table-users 1000001 --blocks 4 | tail -5
python scripts/fullsort_probe.py --users 37888 --reps 3 --path mma --model ComplEx | tail -2
python scripts/fullsort_probe.py --users 37888 --reps 3 --path mma --model TransE --d 100 | tail -2
F="python scripts/fullsort_probe.py --users 37888 --reps 2 --path mma"
$F > gpurun_out/plain_fs_mma.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"fullsort_mma|rescore_topk" -s 2 -c 2 -o gpurun_out/prof_fullsort_mma2 $F > gpurun_out/ncu_fs_mma.log 2>&1
