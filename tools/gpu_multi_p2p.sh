#!/bin/bash
# P2P transport bring-up on N GPUs: slab check (p2p and nccl), then scaling bench for both transports.
set -u
N=${1:-2}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for T in p2p nccl; do
  echo "=== check $T"; F2D_TRANSPORT=$T timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2971$N tests/multi_gpu_check.py > gpurun_out/multi_check_${T}_$N.log 2>&1; echo "exit $?"; grep -E "MULTI_GPU_CHECK|Error|error|timed out" gpurun_out/multi_check_${T}_$N.log | head -5
  echo "=== bench $T"; F2D_TRANSPORT=$T timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2972$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_16384_${N}gpu_$T.log 2>&1; echo "exit $?"; tail -n 1 gpurun_out/bench_16384_${N}gpu_$T.log | cut -c1-200
done
