#!/bin/bash
mkdir -p gpurun_out
python scripts/fullsort_probe.py --users 75776 --reps 8 --path mma --table-users 1000001 --blocks 4 | tail -4
for m in DistMult ComplEx; do
F="python scripts/fullsort_probe.py --users 75776 --reps 3 --path mma --model $m"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fullsort_mma|rescore_topk" -s 4 -c 2 -f -o gpurun_out/prof_fullsort_v4_$m $F > gpurun_out/ncu_fs_$m.log 2>&1; tail -1 gpurun_out/ncu_fs_$m.log
done
