#!/bin/bash
mkdir -p gpurun_out
N=${N:-8}
nvidia-smi -L | wc -l
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "bench rc=$?" >> gpurun_out/bench_n$N.err); tail -3 gpurun_out/bench_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n$N.log').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'value', d['value'], 'ms/step', d['ms_per_step'], 'fwd', d['roofline']['fwd_ms'], 'adam+xchg', d['roofline']['adam_ms'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
for k,v in d['extras'].items(): print(k, v.get('mma'))
PY
