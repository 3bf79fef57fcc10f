#!/bin/bash
# round 2, GPU session M (2 GPUs): exchange tests incl. the owner-sharded optimiser step, bench N=2 (NCCL / owner)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
python -c "import hopwise_b200._abi as a; a.lib(); print('lib ok')"
timeout 600 python -m pytest tests/test_gpu_distributed.py -m gpu -q -x > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_pytest.log
tail -n 5 gpurun_out/r2m_pytest.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
KGE_MULTIMEM=0 timeout 300 $T bench.py --gpus 2 --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2m_bench_n2_nccl.json 2> gpurun_out/r2m_bench_n2_nccl.err; echo "bench n2 nccl rc=$?"
KGE_OWNER_ADAM=1 timeout 300 $T bench.py --gpus 2 --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2m_bench_n2_owner.json 2> gpurun_out/r2m_bench_n2_owner.err; echo "bench n2 owner rc=$?"
python - <<'PY'
import json
for f in ("r2m_bench_n2_nccl", "r2m_bench_n2_owner"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "value", d["value"], "ms", d["ms_per_step"], "fwd", d["roofline"]["fwd_ms"], "adam_ms", d["roofline"]["adam_ms"], "e2e", d["e2e"]["value"], d.get("exchange"), "loss", d.get("final_loss"))
    except Exception as e:
        print(f, "parse failed", e)
PY
tail -n 5 gpurun_out/r2m_bench_n2_owner.err
