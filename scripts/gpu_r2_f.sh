#!/bin/bash
# round 2, GPU session F: whole GPU suite (new tests), smoke, bench.py N=1, launch list + ncu captures for profiles/
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
tail -n 12 gpurun_out/r2f_pytest.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/r2f_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/r2f_smoke.log
timeout 900 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2f_bench.json"))
    print("value", d["value"], "ms", d["ms_per_step"], "median", d["median_ms_per_step"], "e2e", d["e2e"]["value"])
    print("roofline", {k: d["roofline"][k] for k in ("bound", "frac", "dram_frac")}, "hbm", d.get("roofline_hbm", {}).get("frac"))
    for k, v in d.get("extras", {}).items():
        print(k, json.dumps(v)[:700])
    print("cpu", d.get("cpu_baseline"))
except Exception as e:
    print("bench parse failed", e)
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err; echo "ref rc=$?"; cut -c1-600 gpurun_out/r2f_bench_ref.json
