#!/bin/bash
# sweep-kernel duration per build variant (ncu gpu__time_duration of the fullsort_mma / rescore launches)
# usage: [PROBE_ARGS="--model ComplEx"] [TAG=name] bash scripts/gpu_exp_sweep.sh variant...
mkdir -p gpurun_out
F="python scripts/fullsort_probe.py --users 75776 --reps 2 --path mma ${PROBE_ARGS}"
for v in default "$@"; do
  if [ "$v" = default ]; then unset KGE_B200_LIB; else export KGE_B200_LIB=build/variants/libkge_b200_$v.so; fi
  out=gpurun_out/exp_${TAG}$v
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fullsort_mma|rescore_topk" --csv --log-file $out.csv $F > $out.log 2>&1
  echo "== ${TAG}$v: $(grep '_kernel' $out.csv | awk -F'","' '{gsub(/"/,"",$NF); n=split($5,a,"::"); printf "%s=%sus  ", substr(a[n],1,28), $NF/1000}')"
done
