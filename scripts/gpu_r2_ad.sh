#!/bin/bash
# round 2, GPU session AD (1 GPU): resident CTAs of the one-triple-per-warp forward in the HBM regime (cfg5)
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline --workload cfg5_transe_alibaba"
for v in "" m3 m2; do
if [ -n "$v" ]; then export KGE_B200_LIB=build/variants/libkge_b200_$v.so; else unset KGE_B200_LIB; fi
echo "variant=[$v]: $($B | python -c 'import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d["roofline"]; print("ms/step %.4f fwd %.4f adam %.4f frac %.3f" % (d["ms_per_step"], r["fwd_ms"], r["adam_ms"], r["frac"]))')"
done
