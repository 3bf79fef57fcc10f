#!/bin/bash
# round 2, GPU session AJ (1 GPU): launch list of the reference-batch step on the final tree (profiles/)
mkdir -p gpurun_out
B="python bench.py --workload cfg2_transe_ml1m_b2048 --steps 20 --warmup 5 --no-extras --no-cpu-baseline"
$B > gpurun_out/r2aj_plain.log 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_cfg2_b2048.csv $B > /dev/null 2>&1
python - <<'PY'
import csv, collections
lines = [l for l in open("gpurun_out/r2_launches_cfg2_b2048.csv") if l.startswith('"')]
agg = collections.defaultdict(list)
for x in csv.DictReader(lines):
    agg[(x["Kernel Name"][:70], x["Grid Size"])].append(float(x["Metric Value"]))
for k, v in agg.items():
    v = sorted(v)
    print(k, len(v), "median ns", v[len(v) // 2])
PY
