#!/bin/bash
# round 2, GPU session I (2 GPUs): exchange + sharded evaluation tests, smoke (2-GPU branch), bench at N=2 (NCCL and
# the fused in-switch kernel), reference arm under torchrun
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 900 python -m pytest tests/test_gpu_distributed.py tests/test_gpu_scoring.py -m gpu -q > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest.log
tail -n 5 gpurun_out/r2i_pytest.log
timeout 900 python __graft_entry__.py smoke > gpurun_out/r2i_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/r2i_smoke.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $T bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2i_bench_n2.json 2> gpurun_out/r2i_bench_n2.err; echo "bench n2 rc=$?"
KGE_MULTIMEM=1 timeout 900 $T bench.py --gpus 2 --steps 20 --warmup 5 --no-extras > gpurun_out/r2i_bench_n2_mm.json 2> gpurun_out/r2i_bench_n2_mm.err; echo "bench n2 multimem rc=$?"
KGE_MULTIMEM=1 KGE_MULTIMEM_FUSED=0 timeout 900 $T bench.py --gpus 2 --steps 20 --warmup 5 --no-extras > gpurun_out/r2i_bench_n2_mm_unfused.json 2> gpurun_out/r2i_bench_n2_mm_unfused.err; echo "bench n2 multimem unfused rc=$?"
timeout 600 $T bench.py --impl reference --gpus 2 --steps 3 --warmup 3 --no-extras > gpurun_out/r2i_bench_ref_n2.json 2> gpurun_out/r2i_bench_ref_n2.err; echo "ref n2 rc=$?"
python - <<'PY'
import json
for f in ("r2i_bench_n2", "r2i_bench_n2_mm", "r2i_bench_n2_mm_unfused", "r2i_bench_ref_n2"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, "value", d["value"], "ms", d["ms_per_step"], "adam_ms", d.get("roofline", {}).get("adam_ms"), d.get("exchange"), d.get("cpu_baseline", {}).get("cores"))
        for k, v in d.get("extras", {}).items():
            print("   ", k, json.dumps(v)[:500])
    except Exception as e:
        print(f, "parse failed", e)
PY
tail -n 3 gpurun_out/r2i_bench_n2.err
