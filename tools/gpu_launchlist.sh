#!/bin/bash
mkdir -p gpurun_out
F2D_BENCH_SCALING_BASE=0 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 105 -c 80 --csv --log-file gpurun_out/launches_r01_final2.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launches2.log 2>&1
echo "rc=$?"; wc -l gpurun_out/launches_r01_final2.csv; tail -2 gpurun_out/ncu_launches2.log | cut -c1-200
