#!/bin/bash
mkdir -p gpurun_out
F="python scripts/fullsort_probe.py --users 75776 --reps 2 --path mma"
$F > gpurun_out/plain_fs_mma.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"rescore_topk" -s 1 -c 1 -o gpurun_out/prof_rescore $F > gpurun_out/ncu_rescore.log 2>&1
tail -2 gpurun_out/plain_fs_mma.log
