#!/bin/bash
# round-1 GPU session A: correctness, bench, launch list, full ncu capture of the hot kernels
mkdir -p gpurun_out
(timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log); tail -2 gpurun_out/smoke.log
(timeout 1200 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log); tail -15 gpurun_out/pytest.log
(timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err); tail -3 gpurun_out/bench.err; cat gpurun_out/bench.log
B="python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline"
$B > gpurun_out/plain_cfg2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_cfg2.csv $B > gpurun_out/ncu_l2.log 2>&1
$B > gpurun_out/plain_cfg2b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:train_fwd -s 4 -c 1 -o gpurun_out/prof_train_cfg2 $B > gpurun_out/ncu_f2.log 2>&1
$B --workload cfg5_transe_alibaba > gpurun_out/plain_cfg5.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"train_fwd|adam_apply" -s 8 -c 2 -o gpurun_out/prof_train_cfg5 $B --workload cfg5_transe_alibaba > gpurun_out/ncu_f5.log 2>&1
F="python scripts/fullsort_probe.py --users 8192 --reps 2"
$F > gpurun_out/plain_fs.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fullsort_tile -s 1 -c 1 -o gpurun_out/prof_fullsort_distmult $F > gpurun_out/ncu_fs.log 2>&1
cat gpurun_out/plain_fs.log | tail -3
ls -la gpurun_out
