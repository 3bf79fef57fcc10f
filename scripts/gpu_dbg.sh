#!/bin/bash
mkdir -p gpurun_out
echo "== new"; python scripts/dbg_traj.py 2>&1 | tail -40
echo "== base"; KGE_B200_LIB=build/variants/libkge_b200_base.so python scripts/dbg_traj.py 2>&1 | tail -40
