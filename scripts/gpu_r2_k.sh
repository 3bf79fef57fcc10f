#!/bin/bash
# round 2, GPU session K (1 GPU): whole GPU suite (popularity sampler, dynamic negatives, gather kernel, one-call
# step), bench N=1, small-batch launch list + host profile
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log
tail -n 6 gpurun_out/r2k_pytest.log
timeout 600 python bench.py > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2k_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"])
for k in ("cfg2_transe_ml1m_b2048", "cfg5_transe_alibaba_b2048", "cfg2_b2048_device_loader", "cfg1_ml100k_pipeline"):
    print(k, json.dumps(d["extras"].get(k))[:900])
PY
timeout 300 python scripts/host_profile.py > gpurun_out/r2k_host_profile.log 2>&1; head -n 30 gpurun_out/r2k_host_profile.log
B="python bench.py --workload cfg2_transe_ml1m_b2048 --steps 20 --warmup 5 --no-extras --no-cpu-baseline"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2k_launches_b2048.csv $B > /dev/null 2>&1
python - <<'PY'
import csv, collections
lines = [l for l in open("gpurun_out/r2k_launches_b2048.csv") if l.startswith('"')]
agg = collections.defaultdict(list)
for x in csv.DictReader(lines):
    agg[(x["Kernel Name"][:70], x["Grid Size"])].append(float(x["Metric Value"]))
for k, v in agg.items():
    v = sorted(v)
    print(k, len(v), "median ns", v[len(v) // 2])
PY
