#!/bin/bash
CFGS="${XCFGS:-a}" bash scripts/gpu_exp2.sh "$@" | grep -v "^== [A-Za-z0-9]*: $"
(timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err); tail -3 gpurun_out/bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print('train', d['ms_per_step'], d['roofline']['frac'], 'e2e', d['e2e']['ms_per_step'])
for k,v in d['extras'].items():
    if 'mma' in v: print(k, v['mma'])
    else: print(k, {kk:vv for kk,vv in v.items() if kk in ('ms_per_step','hbm_frac','triples_per_s','error')})
PY
