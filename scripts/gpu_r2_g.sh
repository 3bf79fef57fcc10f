#!/bin/bash
# round 2, GPU session G: rescore changes (parity + time), then the ncu evidence for profiles/
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scoring.py tests/test_gpu_sampler_pipeline.py tests/test_gpu_baseline_sizes.py -m gpu -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
tail -n 4 gpurun_out/r2g_pytest.log
TAG=r2g_ bash scripts/gpu_exp_sweep.sh 2>&1 | tee gpurun_out/r2g_exp.log
TAG=r2g_cx_ PROBE_ARGS="--model ComplEx" bash scripts/gpu_exp_sweep.sh 2>&1 | tee -a gpurun_out/r2g_exp.log
TAG=r2g_te_ PROBE_ARGS="--model TransE --d 100" bash scripts/gpu_exp_sweep.sh 2>&1 | tee -a gpurun_out/r2g_exp.log
# launch list of the headline command (time-only pass; per-launch times are cold-cache and serialised)
B="python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline"
$B > gpurun_out/r2g_plain_cfg2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches_cfg2.csv $B > gpurun_out/r2g_ncu_l.log 2>&1
for wl in cfg2_transe_ml1m cfg5_transe_alibaba cfg3_rotate_yelp; do
  $B --workload $wl > gpurun_out/r2g_plain_$wl.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:"train_fwd|adam_apply" -s 8 -c 2 -f -o gpurun_out/r2_prof_train_$wl $B --workload $wl > gpurun_out/r2g_ncu_$wl.log 2>&1
  tail -1 gpurun_out/r2g_ncu_$wl.log
done
# full-sort: launch list of one block + full captures of sweep and rescore for DistMult and ComplEx
for m in DistMult ComplEx; do
  F="python scripts/fullsort_probe.py --users 75776 --reps 3 --path mma --model $m"
  $F > gpurun_out/r2g_plain_fs_$m.log 2>&1 || continue
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_fullsort_$m.csv $F > /dev/null 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fullsort_mma|rescore_topk" -s 2 -c 2 -f -o gpurun_out/r2_prof_fullsort_$m $F > gpurun_out/r2g_ncu_fs_$m.log 2>&1
  tail -1 gpurun_out/r2g_ncu_fs_$m.log
done
