#!/bin/bash
# round 2, GPU session Y (2 GPUs): exchange tests, smoke (2-GPU branch), default bench at N=2 (owner route chosen
# by the multicast probe), reference arm under torchrun
mkdir -p gpurun_out
python -c "import hopwise_b200._abi as a; a.lib(); print('lib ok')"
timeout 600 python -m pytest tests/test_gpu_distributed.py -m gpu -q -x > gpurun_out/r2y_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2y_pytest.log
tail -n 3 gpurun_out/r2y_pytest.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $T bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2y_bench_n2.json 2> gpurun_out/r2y_bench_n2.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
txt = open("gpurun_out/r2y_bench_n2.json").read().strip().splitlines()
print("stdout lines:", len(txt))
d = json.loads(txt[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d.get("exchange"), d.get("rank_split"))
for k, v in d.get("extras", {}).items():
    print("   ", k, json.dumps(v)[:300])
PY
tail -n 3 gpurun_out/r2y_bench_n2.err
timeout 300 $T bench.py --impl reference --gpus 2 --steps 3 --warmup 3 --no-extras | tail -n 1 | cut -c1-300
