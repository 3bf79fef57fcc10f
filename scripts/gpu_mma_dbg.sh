#!/bin/bash
mkdir -p gpurun_out
for f in ${FLAGS:-0 1 2 3}; do
KGE_MMA_DEBUG=$f ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fullsort_mma" --csv --log-file gpurun_out/dbg_$f.csv python scripts/fullsort_probe.py --users ${USERS:-75776} --reps 2 --path mma ${EXTRA} > /dev/null 2>&1
echo "KGE_MMA_DEBUG=$f sweep ns: $(tail -1 gpurun_out/dbg_$f.csv | awk -F'","' '{print $NF}')"
done
