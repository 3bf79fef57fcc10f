#!/bin/bash
# tcgen05 full-sort: parity tests, then per-kernel durations (ncu launch list) per tile shape
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_scoring.py -m gpu -q --timeout 200 -x > gpurun_out/pytest_mma.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_mma.log); tail -8 gpurun_out/pytest_mma.log
for spec in "DistMult 64" "ComplEx 64" "TransE 100" ${EXTRA_SPECS}; do
  set -- $spec
  for cfg in ${CFGS:-auto c}; do
    if [ "$cfg" = auto ]; then unset KGE_MMA_CFG; else export KGE_MMA_CFG=$cfg; fi
    F="python scripts/fullsort_probe.py --users 75776 --reps 3 --path mma --model $1 --d $2"
    timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fullsort_mma|rescore_topk" --csv --log-file gpurun_out/l_$1_$cfg.csv $F > gpurun_out/l_$1_$cfg.log 2>&1
    echo "== $1 d=$2 cfg=$cfg: $(grep gpu__time_duration gpurun_out/l_$1_$cfg.csv | tail -2 | awk -F'","' '{printf "%s %s ns   ", substr($5,17,24), $NF}') | $(tail -1 gpurun_out/l_$1_$cfg.log | cut -c1-60)"
  done
done
