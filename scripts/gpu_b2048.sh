#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --workload cfg2_transe_ml1m_b2048 --steps 20 --warmup 5 --no-extras --no-cpu-baseline"
$B | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("ms/step", d["ms_per_step"], "fwd", d["roofline"]["fwd_ms"], "adam", d["roofline"]["adam_ms"], "e2e", d["e2e"]["ms_per_step"])'
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_b2048.csv $B > /dev/null 2>&1
grep gpu__time_duration gpurun_out/launches_b2048.csv | awk -F'","' '{printf "%-60s %s\n", substr($5,1,60), $NF}' | sed -n 60,80p
