#!/bin/bash
# round 2, GPU session J2 (8 GPUs): the driver's own command at N=8 (default routes, extras included)
mkdir -p gpurun_out
python -c "import hopwise_b200._abi as a; a.lib(); print('lib ok')"
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
timeout 420 $T bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2j2_bench_n8.json 2> gpurun_out/r2j2_bench_n8.err; echo "bench n8 rc=$?"
python - <<'PY'
import json
txt = open("gpurun_out/r2j2_bench_n8.json").read().strip().splitlines()
print("stdout lines:", len(txt))
d = json.loads(txt[-1])
print("value", d["value"], "ms", d["ms_per_step"], "fwd", d["roofline"]["fwd_ms"], "adam", d["roofline"]["adam_ms"], "e2e", d["e2e"]["value"], d.get("exchange", "")[:30], d.get("rank_split"), "loss", d["final_loss"])
for k, v in d.get("extras", {}).items():
    print("   ", k, json.dumps(v)[:260])
PY
tail -n 3 gpurun_out/r2j2_bench_n8.err
