#!/bin/bash
# round 2, GPU session E: deferred appends; whole GPU test suite incl. the hopwise pipeline test; bench.py
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
tail -n 6 gpurun_out/r2e_pytest.log
TAG=r2e_ bash scripts/gpu_exp_sweep.sh NMMA2 2>&1 | tee gpurun_out/r2e_exp.log
timeout 900 python bench.py > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r2e_bench.err | tail -n 5
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2e_bench.json"))
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"])
    print("roofline", {k: d["roofline"][k] for k in ("bound", "frac", "dram_frac")}, "hbm", d.get("roofline_hbm", {}).get("frac"))
    for k, v in d.get("extras", {}).items():
        print(k, json.dumps(v)[:400])
    print("cpu", d.get("cpu_baseline"))
except Exception as e:
    print("bench parse failed", e)
PY
bash scripts/gpu_train_variants.sh TPW TPW2 MC3 2>&1 | tee gpurun_out/r2e_train_variants.log
