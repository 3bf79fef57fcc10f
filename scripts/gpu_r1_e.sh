#!/bin/bash
# round-1 GPU session E: full GPU test run, bench, ncu captures of the train step on three workloads
mkdir -p gpurun_out
(timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log); tail -2 gpurun_out/smoke.log
(timeout 1200 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log); tail -5 gpurun_out/pytest.log
(timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err); tail -3 gpurun_out/bench.err; cat gpurun_out/bench.log
B="python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline"
for w in cfg2_transe_ml1m cfg5_transe_alibaba cfg3_rotate_yelp; do
$B --workload $w > gpurun_out/plain_$w.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"train_fwd|adam_apply" -s 8 -c 2 -o gpurun_out/prof_train_$w $B --workload $w > gpurun_out/ncu_$w.log 2>&1
done
ls -la gpurun_out | tail -8
