#!/bin/bash
# 1-GPU box: what the driver runs at round end, in order
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== tests"; timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/tests.log 2>&1; echo "exit $?"; tail -n 3 gpurun_out/tests.log | cut -c1-200
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "=== ref arm"; timeout 600 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -n 1 gpurun_out/bench_ref.log | cut -c1-160
echo "=== bench 1"; timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench_final_1.log 2>&1; tail -n 1 gpurun_out/bench_final_1.log | cut -c1-200
