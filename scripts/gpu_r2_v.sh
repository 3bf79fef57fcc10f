#!/bin/bash
# round 2, GPU session V (1 GPU): TransH golden trajectory after the factor fix, pipelined e2e loop, loader probe
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_scoring.py -m gpu -q 2>&1 | tail -n 3
timeout 600 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2v_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", json.dumps(d["e2e"])[:700])
PY
timeout 300 python scripts/loader_probe.py 2>&1 | tail -n 8
