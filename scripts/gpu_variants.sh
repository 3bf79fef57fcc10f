#!/bin/bash
for v in "$@"; do
echo "== variant $v"
KGE_B200_LIB=build/variants/libkge_b200_$v.so timeout 120 python scripts/fullsort_probe.py --users 75776 --reps 4 --path mma | tail -2
done
echo "== default"; timeout 120 python scripts/fullsort_probe.py --users 75776 --reps 4 --path mma | tail -2
