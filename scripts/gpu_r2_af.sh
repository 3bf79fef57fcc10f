#!/bin/bash
# round 2, GPU session AF (1 GPU): whole GPU suite, smoke, bench N=1 and the reference arm on the final tree
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2af_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2af_pytest.log
grep -E "^E  |passed|failed|FAILED|rc=" gpurun_out/r2af_pytest.log | head -n 20 | cut -c1-300
timeout 900 python __graft_entry__.py smoke > gpurun_out/r2af_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/r2af_smoke.log
timeout 600 python bench.py > gpurun_out/r2af_bench.json 2> gpurun_out/r2af_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2af_bench.json").read().strip().splitlines()[-1])
e = d["e2e"]
print("value", d["value"], "ms", d["ms_per_step"], "e2e", e["value"], e["ms_per_step"], "ref-order", e["reference_loop_order_ms_per_step"])
for k in ("cfg4_distmult", "cfg4_complex"):
    m = d["extras"][k]["mma"]; print(k, m["users_per_s"], m["mean_ms_per_block"], m["median_ms_per_block"])
print(json.dumps(d["extras"]["cfg2_b2048_device_loader"]))
PY
