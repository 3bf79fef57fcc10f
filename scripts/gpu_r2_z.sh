#!/bin/bash
# round 2, GPU session Z (1 GPU): loader fast path tests, bench N=1, then the ncu evidence for profiles/ on the final
# kernels (summaries produced on the box)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_loader.py tests/test_gpu_sampler.py -m gpu -q 2>&1 | tail -n 3
timeout 600 python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2z_bench.json").read().strip().splitlines()[-1])
e = d["e2e"]
print("value", d["value"], "ms", d["ms_per_step"], "e2e", e["value"], e["ms_per_step"], "ref-order", e["reference_loop_order_ms_per_step"])
print(json.dumps(d["extras"]["cfg2_b2048_device_loader"]))
PY
B="python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches_cfg2.csv $B > gpurun_out/r2z_ncu_l.log 2>&1
for wl in cfg2_transe_ml1m cfg5_transe_alibaba cfg3_rotate_yelp; do
  ncu --set full --clock-control none --import-source on -k regex:"train_fwd|adam_apply" -s 8 -c 2 -f -o /tmp/r2_prof_train_$wl $B --workload $wl > gpurun_out/r2z_ncu_$wl.log 2>&1
  python scripts/ncu_summary.py /tmp/r2_prof_train_$wl.ncu-rep gpurun_out/r2_train_$wl.txt --top 14 | tail -1
done
du -sh gpurun_out
