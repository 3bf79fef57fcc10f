#!/bin/bash
mkdir -p gpurun_out
F="python scripts/fullsort_probe.py --users 75776 --reps 4 --path mma ${PROBE_ARGS}"
$F | tail -2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_fs_all.csv $F > /dev/null 2>&1
grep gpu__time_duration gpurun_out/launches_fs_all.csv | awk -F'","' '{printf "%-70s %s\n", substr($5,1,70), $NF}' | tail -14
