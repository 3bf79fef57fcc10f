#!/bin/bash
# ncu captures of the train step (launch list of the headline command + one --set full capture per workload)
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline"
$B > gpurun_out/plain_cfg2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_cfg2.csv $B > gpurun_out/ncu_l2.log 2>&1
for wl in cfg2_transe_ml1m cfg5_transe_alibaba cfg3_rotate_yelp; do
  $B --workload $wl > gpurun_out/plain_$wl.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:"train_fwd|adam_apply" -s 8 -c 2 -f -o gpurun_out/prof_train_$wl $B --workload $wl > gpurun_out/ncu_$wl.log 2>&1
  tail -1 gpurun_out/ncu_$wl.log
done
