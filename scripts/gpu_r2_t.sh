#!/bin/bash
mkdir -p gpurun_out
python scripts/debug_transh.py 2>&1 | tail -n 12
