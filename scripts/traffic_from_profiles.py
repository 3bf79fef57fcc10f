#!/usr/bin/env python
"""profiles/r2_train_<workload>.txt (ncu summaries) -> profiles/r2_traffic.json: DRAM bytes (read + write) per
launch of the two kernels of one train step, per workload.  bench.py reports them as `roofline.traffic`."""
import glob
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out = {}
for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r2_train_*.txt"))):
    name = os.path.basename(path)[len("r2_train_"):-4]
    kernels, cur = {}, None
    for line in open(path):
        m = re.match(r"## void <unnamed>::(\w+)<", line)
        if m:
            cur = m.group(1)
            kernels.setdefault(cur, 0.0)
            continue
        if line.startswith("## hottest"):
            cur = None
        m = re.match(r"\s+dram (read|write)\s+([0-9.]+) (\w+)$", line.rstrip())
        if m and cur:
            kernels[cur] += float(m.group(2)) * UNIT[m.group(3)]
    if kernels:
        out[name] = {"per_launch_bytes": kernels, "per_step_bytes": sum(kernels.values()), "source": os.path.basename(path)}
json.dump(out, open(os.path.join(ROOT, "profiles", "r2_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
