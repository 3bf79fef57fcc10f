#!/bin/bash
# round 2, GPU session H: ncu evidence for profiles/ -- summaries are produced on the box (the .ncu-rep files are
# too large to travel together); only text / csv comes back, plus the DistMult sweep report for the source page
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline"
$B > gpurun_out/r2h_plain_cfg2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches_cfg2.csv $B > gpurun_out/r2h_ncu_l.log 2>&1
for wl in cfg2_transe_ml1m cfg5_transe_alibaba cfg3_rotate_yelp; do
  $B --workload $wl > gpurun_out/r2h_plain_$wl.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:"train_fwd|adam_apply" -s 8 -c 2 -f -o /tmp/r2_prof_train_$wl $B --workload $wl > gpurun_out/r2h_ncu_$wl.log 2>&1
  python scripts/ncu_summary.py /tmp/r2_prof_train_$wl.ncu-rep gpurun_out/r2_train_$wl.txt --top 14 | tail -1
done
for m in DistMult ComplEx; do
  F="python scripts/fullsort_probe.py --users 75776 --reps 3 --path mma --model $m"
  $F > gpurun_out/r2h_plain_fs_$m.log 2>&1 || continue
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_fullsort_$m.csv $F > /dev/null 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fullsort_mma|rescore_topk" -s 2 -c 2 -f -o /tmp/r2_prof_fullsort_$m $F > gpurun_out/r2h_ncu_fs_$m.log 2>&1
  python scripts/ncu_summary.py /tmp/r2_prof_fullsort_$m.ncu-rep gpurun_out/r2_fullsort_$m.txt --top 24 | tail -1
done
cp /tmp/r2_prof_fullsort_DistMult.ncu-rep gpurun_out/ 2>/dev/null
du -sh gpurun_out
