#!/bin/bash
for cfg in ${CFGS:-f a}; do
export KGE_MMA_CFG=$cfg
echo "#### cfg $cfg"
bash scripts/gpu_exp_sweep.sh "$@" | grep -v "^$"
cd gpurun_out; for v in default "$@"; do echo "== $v: $(grep "gpu__time_duration" exp_$v.csv | awk -F'","' '{printf "%s %s  ", substr($5,17,24), $NF}' | tail -1)"; done; cd ..
done
