#!/bin/bash
# round 2, GPU session L (1 GPU): whole GPU suite, bench N=1, small-batch kernel durations
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_pytest.log
tail -n 8 gpurun_out/r2l_pytest.log
timeout 600 python bench.py > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2l_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "fwd", d["roofline"]["fwd_ms"], "adam", d["roofline"]["adam_ms"], "e2e", d["e2e"]["value"])
for k in ("cfg2_transe_ml1m_b2048", "cfg5_transe_alibaba", "cfg5_transe_alibaba_b2048", "cfg3_rotate_yelp", "cfg2_b2048_device_loader", "cfg1_ml100k_pipeline"):
    print(k, json.dumps(d["extras"].get(k))[:1000])
PY
B="python bench.py --workload cfg2_transe_ml1m_b2048 --steps 20 --warmup 5 --no-extras --no-cpu-baseline"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2l_launches_b2048.csv $B > /dev/null 2>&1
python - <<'PY'
import csv, collections
lines = [l for l in open("gpurun_out/r2l_launches_b2048.csv") if l.startswith('"')]
agg = collections.defaultdict(list)
for x in csv.DictReader(lines):
    agg[(x["Kernel Name"][:70], x["Grid Size"])].append(float(x["Metric Value"]))
for k, v in agg.items():
    v = sorted(v)
    print(k, len(v), "median ns", v[len(v) // 2])
PY
