#!/bin/bash
# N-GPU bench of the train step with the dense route reduced in the switch (multimem) and through NCCL
mkdir -p gpurun_out
N=${N:-8}
for mm in 1 0; do
(KGE_MULTIMEM=$mm timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2955$mm bench.py --gpus $N --steps 10 --warmup 3 --no-extras > gpurun_out/bench_n${N}_mm$mm.log 2> gpurun_out/bench_n${N}_mm$mm.err; echo "rc=$?" >> gpurun_out/bench_n${N}_mm$mm.err)
tail -1 gpurun_out/bench_n${N}_mm$mm.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_n${N}_mm$mm.log').read().strip().splitlines()[-1])
print('multimem=$mm N', d['n_gpus'], 'value %.3e' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'fwd %.4f' % d['roofline']['fwd_ms'], 'adam+xchg %.4f' % d['roofline']['adam_ms'], 'e2e ms %.4f' % d['e2e']['ms_per_step'], 'launches', d['gpu_launches'])
"
done
