#!/bin/bash
# 8-GPU box: slab check at 8 ranks, scaling bench at 8, 4 and 2 ranks (16384^2, K=80).
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/multi8_gpus.txt 2>&1
echo "=== check 8"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 tests/multi_gpu_check.py > gpurun_out/multi_check_8.log 2>&1; echo "exit $?"; tail -n 3 gpurun_out/multi_check_8.log
for N in 8 4 2; do
  echo "=== bench $N GPUs"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2963$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_16384_${N}gpu.log 2>&1; echo "exit $?"; tail -n 1 gpurun_out/bench_16384_${N}gpu.log | cut -c1-400
done
echo "=== bench 1 GPU 16384"; timeout 900 python bench.py --gpus 1 --size 16384 --steps 5 --warmup 3 > gpurun_out/bench_16384_1gpu.log 2>&1; echo "exit $?"; tail -n 1 gpurun_out/bench_16384_1gpu.log | cut -c1-300
echo "=== NCCL_DEBUG probe"; NCCL_DEBUG=INFO timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 2 --steps 2 --warmup 3 2>&1 | grep -iE "NVLS|P2P|via|channel" | head -20 > gpurun_out/nccl_info.txt; wc -l gpurun_out/nccl_info.txt
