#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cpu_semantics.py -x -q > gpurun_out/test_cpu_sem.log 2>&1
echo "pytest rc=$?" >> gpurun_out/test_cpu_sem.log
tail -3 gpurun_out/test_cpu_sem.log
for cfg in "256 20 20" "1024 20 5" "1024 80 3" "4096 20 3"; do
  timeout 120 python tools/gs_bench.py $cfg >> gpurun_out/gs_bench.log 2>&1
done
cat gpurun_out/gs_bench.log
for cfg in "256 20" "4096 20"; do
  F2D_LIB_PATH=fluid-2d_b200/libf2d_gstime.so timeout 120 python tools/gs_timing.py $cfg >> gpurun_out/gs_timing.log 2>&1
done
cat gpurun_out/gs_timing.log
