#!/bin/bash
# one ncu --set full capture of the sweep kernel (after a plain run exits 0)
mkdir -p gpurun_out
F="python scripts/fullsort_probe.py --users 75776 --reps 2 --path mma ${PROBE_ARGS}"
$F > gpurun_out/plain_sweep.log 2>&1 || { tail -5 gpurun_out/plain_sweep.log; exit 1; }
tail -1 gpurun_out/plain_sweep.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fullsort_mma" -s 1 -c 1 -f -o gpurun_out/${OUT:-prof_sweep} $F > gpurun_out/ncu_sweep.log 2>&1
tail -2 gpurun_out/ncu_sweep.log
