#!/bin/bash
# two-GPU session: NCCL row-sparse step parity, then the bench at N=2
mkdir -p gpurun_out
nvidia-smi -L
(timeout 600 python -m pytest tests/test_gpu_distributed.py -m gpu -q --timeout 200 -x > gpurun_out/pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_2gpu.log); tail -15 gpurun_out/pytest_2gpu.log
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err; echo "bench rc=$?" >> gpurun_out/bench_n2.err); tail -5 gpurun_out/bench_n2.err; cat gpurun_out/bench_n2.log
(KGE_MULTIMEM=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-extras > gpurun_out/bench_n2_nccl.log 2> gpurun_out/bench_n2_nccl.err; echo "bench(nccl) rc=$?" >> gpurun_out/bench_n2_nccl.err); tail -2 gpurun_out/bench_n2_nccl.err; cut -c1-330 gpurun_out/bench_n2_nccl.log
