#!/bin/bash
# round 2, GPU session AE (2 GPUs): default bench at N=2 with extras (GC off in timed regions: does the slow second
# full-sort block go away?)
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $T bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2ae_bench_n2.json 2> gpurun_out/r2ae_bench_n2.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2ae_bench_n2.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d.get("rank_split"))
for k in ("cfg4_distmult", "cfg4_complex"):
    m = d["extras"][k]["mma"]
    print(k, "mean", m["mean_ms_per_block"], "median", m["median_ms_per_block"], m["per_block_ms"])
print(d["extras"]["cfg4_distmult_full_eval"]["ms"])
PY
tail -n 2 gpurun_out/r2ae_bench_n2.err
