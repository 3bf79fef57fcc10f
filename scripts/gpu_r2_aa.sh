#!/bin/bash
# round 2, GPU session AA (2 GPUs): owner-sharded Adam with four quads in flight -- parity tests, N=2 step time
mkdir -p gpurun_out
python -c "import hopwise_b200._abi as a; a.lib(); print('lib ok')"
timeout 600 python -m pytest tests/test_gpu_distributed.py -m gpu -q -x -k "owner or multimem or True" > gpurun_out/r2aa_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2aa_pytest.log
tail -n 3 gpurun_out/r2aa_pytest.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $T bench.py --gpus 2 --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2aa_bench_n2.json 2> gpurun_out/r2aa_bench_n2.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2aa_bench_n2.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "fwd", d["roofline"]["fwd_ms"], "adam", d["roofline"]["adam_ms"], "e2e", d["e2e"]["value"], d.get("rank_split"), "loss", d["final_loss"])
PY
tail -n 3 gpurun_out/r2aa_bench_n2.err
