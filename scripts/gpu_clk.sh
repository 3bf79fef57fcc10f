#!/bin/bash
for cfg in f a; do echo "### cfg $cfg"; KGE_MMA_CFG=$cfg KGE_B200_LIB=build/variants/libkge_b200_CLK.so python scripts/fullsort_probe.py --users 75776 --reps 1 --path mma 2>&1 | grep -v fallback | sort | head -20; done
