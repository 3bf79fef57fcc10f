#!/bin/bash
N=${N:-8}
for algo in default NVLS Ring Tree; do
  if [ "$algo" = default ]; then unset NCCL_ALGO; else export NCCL_ALGO=$algo; fi
  timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 scripts/allreduce_probe.py 2>&1 | grep "all_reduce"
done
unset NCCL_ALGO
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,TUNING timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 scripts/allreduce_probe.py 2>&1 | grep -i "nvls\|algo\|multicast" | head -8
