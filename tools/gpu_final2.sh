#!/bin/bash
# 2-GPU box: everything the driver runs at round end, in order (auto transport).
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== tests"; timeout 1200 python -m pytest tests -q -m gpu --maxfail=10 > gpurun_out/tests.log 2>&1; echo "exit $?"; tail -n 4 gpurun_out/tests.log | cut -c1-200
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "=== ref arm"; timeout 600 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 | cut -c1-200
echo "=== bench 1"; timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench_final_1.log 2>&1; tail -n 1 gpurun_out/bench_final_1.log | cut -c1-250
echo "=== bench 2 (auto transport)"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29901 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_final_2.log 2>&1; echo "exit $?"; tail -n 1 gpurun_out/bench_final_2.log | cut -c1-250; grep -o '"transport": "[a-z0-9]*"' gpurun_out/bench_final_2.log
echo "=== ref arm under torchrun"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29902 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 2>&1 | grep '^{' | cut -c1-150
