#!/bin/bash
# 2-GPU box: slab check (p2p + nccl) and the 2-GPU bench after the scratch-pool change
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== slab check (auto transport)"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29911 tests/multi_gpu_check.py > gpurun_out/multi_check_2.log 2>&1; echo "exit $?"; tail -n 3 gpurun_out/multi_check_2.log | cut -c1-200
echo "=== slab check (nccl)"; F2D_TRANSPORT=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29912 tests/multi_gpu_check.py > gpurun_out/multi_check_nccl_2.log 2>&1; echo "exit $?"; tail -n 2 gpurun_out/multi_check_nccl_2.log | cut -c1-200
echo "=== bench 2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29913 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_final_2.log 2>&1; echo "exit $?"; tail -n 1 gpurun_out/bench_final_2.log | cut -c1-300
