#!/bin/bash
# 1-GPU box: everything the driver runs at round end (tests, smoke, both bench arms) + ncu capture of the GS kernel
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== tests"; timeout 1500 python -m pytest tests -q -m gpu --maxfail=10 > gpurun_out/tests.log 2>&1; echo "exit $?"; tail -n 4 gpurun_out/tests.log | cut -c1-200
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "=== ref arm"; timeout 600 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 | cut -c1-200
echo "=== bench 1"; timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench_final_1.log 2>&1; tail -n 1 gpurun_out/bench_final_1.log | cut -c1-300
echo "=== ncu gs"; timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gs_relax -s 1 -c 1 -f -o gpurun_out/gs_relax_r01_final python tools/gs_bench.py 1024 20 1 > gpurun_out/ncu_gs.log 2>&1; echo "exit $?"
