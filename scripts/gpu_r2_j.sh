#!/bin/bash
# round 2, GPU session J (8 GPUs): headline step at N=8 -- NCCL all-reduce, in-switch all-reduce with in-kernel
# barriers, owner-sharded Adam over the switch
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -n 8
python -c "import hopwise_b200._abi as a; a.lib(); print('lib ok')"
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
A="bench.py --gpus 8 --steps 20 --warmup 5 --no-extras --no-cpu-baseline"
KGE_OWNER_ADAM=1 timeout 240 $T $A > gpurun_out/r2j_bench_n8_owner.json 2> gpurun_out/r2j_bench_n8_owner.err; echo "owner rc=$?"
KGE_MULTIMEM=0 timeout 240 $T $A > gpurun_out/r2j_bench_n8_nccl.json 2> gpurun_out/r2j_bench_n8_nccl.err; echo "nccl rc=$?"
KGE_MULTIMEM=1 timeout 240 $T $A > gpurun_out/r2j_bench_n8_mm.json 2> gpurun_out/r2j_bench_n8_mm.err; echo "multimem fused rc=$?"
python - <<'PY'
import json
for f in ("r2j_bench_n8_owner", "r2j_bench_n8_nccl", "r2j_bench_n8_mm"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "value", d["value"], "ms", d["ms_per_step"], "fwd", d["roofline"]["fwd_ms"], "adam_ms", d["roofline"]["adam_ms"], "e2e", d["e2e"]["value"], "loss", d.get("final_loss"), d.get("exchange", "")[:40])
    except Exception as e:
        print(f, "parse failed", e)
PY
tail -n 3 gpurun_out/r2j_bench_n8_owner.err
