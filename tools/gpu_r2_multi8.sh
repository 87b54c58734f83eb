#!/bin/bash
# round 2, 8-GPU call: slab == single GPU (step, pipelined solve(), CFL guard), the published 16384^2 K=80 step on 8 slabs
# against the live reference GPU solver, then the 16384^2 bench line with its own 1-GPU base
set -u
N=${1:-8}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T=p2p
echo "=== check $T"; F2D_CHECK_BIG=16384 F2D_TRANSPORT=$T timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2971$N tests/multi_gpu_check.py > gpurun_out/multi_check_${T}_$N.log 2>&1; echo "exit $?"; grep -E "MULTI_GPU_CHECK|Error|error|timed out|\"case\"|\"ok\"" gpurun_out/multi_check_${T}_$N.log | paste - - | head -24
echo "=== bench $T"; F2D_TRANSPORT=$T timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2972$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_16384_${N}gpu_$T.log 2>&1; echo "exit $?"; tail -n 1 gpurun_out/bench_16384_${N}gpu_$T.log | cut -c1-300
