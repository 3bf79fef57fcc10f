#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do timeout 300 python __graft_entry__.py smoke > gpurun_out/r2am_smoke_$i.log 2>&1; echo "smoke $i rc=$?"; tail -n 1 gpurun_out/r2am_smoke_$i.log | cut -c1-200; done
