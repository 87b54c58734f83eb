#!/bin/bash
# Multi-GPU session: slab self-consistency check + scaling bench.  Usage: gpu_multi.sh NGPUS
set -u
N=${1:-2}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/multi_gpus.txt 2>&1
nvidia-smi topo -m >> gpurun_out/multi_gpus.txt 2>&1
echo "=== check"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tests/multi_gpu_check.py > gpurun_out/multi_check_$N.log 2>&1; echo "exit $?"; tail -n 30 gpurun_out/multi_check_$N.log
echo "=== bench 1 GPU at 16384"; timeout 900 python bench.py --gpus 1 --size 16384 --steps 5 --warmup 3 > gpurun_out/bench_16384_1.log 2>&1; echo "exit $?"; tail -n 3 gpurun_out/bench_16384_1.log
echo "=== bench $N GPUs"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_16384_$N.log 2>&1; echo "exit $?"; tail -n 5 gpurun_out/bench_16384_$N.log
