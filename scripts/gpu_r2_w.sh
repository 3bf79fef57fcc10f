#!/bin/bash
# round 2, GPU session W (1 GPU): loader prefetch parity, bench with the pipelined e2e loop and the loader leg, probe
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_loader.py tests/test_gpu_sampler.py -m gpu -q 2>&1 | tail -n 3
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2w_bench.json").read().strip().splitlines()[-1])
e = d["e2e"]
print("value", d["value"], "ms", d["ms_per_step"], "e2e", e["value"], e["ms_per_step"], "median", e["median_ms_per_step"], "ref-order", e["reference_loop_order_ms_per_step"])
print(json.dumps(d["extras"]["cfg2_b2048_device_loader"]))
print(json.dumps(d["extras"]["cfg2_transe_ml1m_b2048"])[:600])
PY
timeout 300 python scripts/loader_probe.py 2>&1 | tail -n 20
