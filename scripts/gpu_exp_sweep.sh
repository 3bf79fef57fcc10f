#!/bin/bash
# sweep-kernel duration per build variant (ncu gpu__time_duration of the last fullsort_mma launch)
mkdir -p gpurun_out
F="python scripts/fullsort_probe.py --users 75776 --reps 2 --path mma ${PROBE_ARGS}"
for v in default "$@"; do
  if [ "$v" = default ]; then unset KGE_B200_LIB; else export KGE_B200_LIB=build/variants/libkge_b200_$v.so; fi
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fullsort_mma|rescore_topk" --csv --log-file gpurun_out/exp_$v.csv $F > gpurun_out/exp_$v.log 2>&1
  echo "== $v: $(grep -o '"[a-z_]*_kernel[^"]*","[0-9]*","gpu__time_duration.sum","[a-z]*","[0-9.,]*"' gpurun_out/exp_$v.csv | sed 's/void <unnamed>:://; s/(<unnamed>::[A-Za-z]*)//' | awk -F'","' '{printf "%s=%s%s  ", substr($1,2,22), $5, $4}')"
done
