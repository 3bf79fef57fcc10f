#!/bin/bash
B="python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline"
for v in "" "$@"; do
for w in cfg2_transe_ml1m cfg5_transe_alibaba cfg3_rotate_yelp; do
if [ -n "$v" ]; then export KGE_B200_LIB=build/variants/libkge_b200_$v.so; else unset KGE_B200_LIB; fi
echo "variant=[$v] $w: $($B --workload $w | python -c 'import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d["roofline"]; print("ms/step %.4f fwd %.4f adam %.4f frac %.3f e2e_ms %.4f" % (d["ms_per_step"], r["fwd_ms"], r["adam_ms"], r["frac"], d["e2e"]["ms_per_step"]))')"
done; done
