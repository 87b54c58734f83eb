#!/bin/bash
# GPU run: parity of the fluid_solver_cpu-compatible mode + first timings (bounded: every command under timeout)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/gpu_gs.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_cpu_semantics.py -x -q > gpurun_out/test_cpu_sem.log 2>&1
echo "pytest rc=$?" >> gpurun_out/test_cpu_sem.log
tail -5 gpurun_out/test_cpu_sem.log
for cfg in "256 20 20" "1024 20 5" "1024 80 3" "4096 20 3"; do
  timeout 120 python tools/gs_bench.py $cfg >> gpurun_out/gs_bench.log 2>&1
done
cat gpurun_out/gs_bench.log
