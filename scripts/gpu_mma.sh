#!/bin/bash
# tcgen05 full-sort: parity tests, then timings of the sweep + rescore on the config-4 shape
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_scoring.py -m gpu -q --timeout 120 -k "mma" -x > gpurun_out/pytest_mma.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_mma.log); tail -15 gpurun_out/pytest_mma.log
for tn in ${TNS:-64}; do
echo "== TN=$tn"
KGE_MMA_TN=$tn timeout 120 python scripts/fullsort_probe.py --users 37888 --reps 3 --path mma | tail -1
KGE_MMA_TN=$tn timeout 120 python scripts/fullsort_probe.py --users 75776 --reps 4 --path mma | tail -2
KGE_MMA_TN=$tn timeout 120 python scripts/fullsort_probe.py --users 75776 --reps 3 --path mma --model ComplEx | tail -1
KGE_MMA_TN=$tn timeout 120 python scripts/fullsort_probe.py --users 75776 --reps 3 --path mma --model TransE --d 100 | tail -1
done
F="python scripts/fullsort_probe.py --users 75776 --reps 2 --path mma"
$F > gpurun_out/plain_fs_mma.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fullsort_mma|rescore_topk" --csv --log-file gpurun_out/launches_fs_mma.csv $F > /dev/null 2>&1
grep -o '"[a-z_:<>(), A-Za-z0-9]*","[0-9]*","gpu__time_duration.sum","[a-z]*","[0-9.,]*"' gpurun_out/launches_fs_mma.csv | tail -6 || tail -5 gpurun_out/launches_fs_mma.csv
if [ -n "$FULL" ]; then
ncu --set full --clock-control none --import-source on -k regex:"fullsort_mma|rescore_topk" -s 2 -c 2 -o gpurun_out/prof_fullsort_mma3 $F > gpurun_out/ncu_fs_mma.log 2>&1
fi
