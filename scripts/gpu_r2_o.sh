#!/bin/bash
# round 2, GPU session O (1 GPU): whole GPU suite + smoke after the optimiser-arithmetic change, bench N=1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2o_pytest.log
tail -n 6 gpurun_out/r2o_pytest.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/r2o_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/r2o_smoke.log
timeout 600 python bench.py > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2o_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "fwd", d["roofline"]["fwd_ms"], "adam", d["roofline"]["adam_ms"], "e2e", d["e2e"]["value"])
for k in ("cfg2_transe_ml1m_b2048", "cfg5_transe_alibaba", "cfg5_transe_alibaba_b2048", "cfg3_rotate_yelp", "cfg2_b2048_device_loader"):
    print(k, json.dumps(d["extras"].get(k))[:1000])
PY
