#!/bin/bash
CFGS="auto" bash scripts/gpu_mma2.sh
CFGS="${XCFGS:-a}" bash scripts/gpu_exp2.sh "$@" | grep -v "^== [A-Za-z0-9]*: $"
