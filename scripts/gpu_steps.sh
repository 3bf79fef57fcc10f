#!/bin/bash
mkdir -p gpurun_out
python scripts/steps_probe.py
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 9000 --launch-count 24 --csv --log-file gpurun_out/launches_steps.csv python scripts/steps_probe.py > /dev/null 2>&1
grep gpu__time_duration gpurun_out/launches_steps.csv | awk -F'","' '{printf "%-60s %s\n", substr($5,1,60), $NF}' | tail -12
