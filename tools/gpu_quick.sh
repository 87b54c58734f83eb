#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== tests"; timeout 1200 python -m pytest tests -q -m gpu --maxfail=20 > gpurun_out/tests.log 2>&1; echo "exit $?"; tail -n 25 gpurun_out/tests.log | cut -c1-300
echo "=== bench"; timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; tail -n 1 gpurun_out/bench.log | cut -c1-330
echo "=== bench unfused src"; F2D_FUSE_SOURCES=0 timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_unfused.log 2>&1; tail -n 1 gpurun_out/bench_unfused.log | cut -c1-330
echo "=== bench 16384"; timeout 600 python bench.py --size 16384 --steps 5 --warmup 3 > gpurun_out/bench16384.log 2>&1; tail -n 1 gpurun_out/bench16384.log | cut -c1-330
