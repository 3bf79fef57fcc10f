#!/bin/bash
# round 2, GPU session X (1 GPU): whole GPU suite, smoke, bench N=1 (loader leg with the late loss read)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2x_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2x_pytest.log
grep -E "^E  |passed|failed|FAILED|rc=" gpurun_out/r2x_pytest.log | head -n 20 | cut -c1-300
timeout 600 python bench.py > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2x_bench.json").read().strip().splitlines()[-1])
e = d["e2e"]
print("value", d["value"], "ms", d["ms_per_step"], "e2e", e["value"], e["ms_per_step"], "median", e["median_ms_per_step"], "ref-order", e["reference_loop_order_ms_per_step"])
for k in ("cfg2_b2048_device_loader", "cfg2_transe_ml1m_b2048", "cfg5_transe_alibaba_b2048"):
    print(k, json.dumps(d["extras"][k])[:700])
PY
