#!/bin/bash
# round 2, GPU session AG (8-GPU box): N=4 headline (owner route), then the driver's command at N=8 on the final tree
mkdir -p gpurun_out
python -c "import hopwise_b200._abi as a; a.lib(); print('lib ok')"
T4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531"
timeout 240 $T4 bench.py --gpus 4 --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2ag_bench_n4.json 2> gpurun_out/r2ag_bench_n4.err; echo "bench n4 rc=$?"
T8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
timeout 420 $T8 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2ag_bench_n8.json 2> gpurun_out/r2ag_bench_n8.err; echo "bench n8 rc=$?"
python - <<'PY'
import json
for f in ("r2ag_bench_n4", "r2ag_bench_n8"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d.get("exchange", "")[:24], d.get("rank_split"), "loss", d["final_loss"])
        for k in ("cfg4_distmult", "cfg4_complex"):
            if k in d.get("extras", {}):
                m = d["extras"][k]["mma"]; print("   ", k, m["users_per_s"], m["mean_ms_per_block"], m["median_ms_per_block"])
        if "cfg4_distmult_full_eval" in d.get("extras", {}):
            print("    full eval ms", d["extras"]["cfg4_distmult_full_eval"]["ms"])
    except Exception as e:
        print(f, "parse failed", e)
PY
tail -n 2 gpurun_out/r2ag_bench_n8.err
