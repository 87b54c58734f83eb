#!/bin/bash
# A/B of the GS kernel's register budget / residency (4, 5, 6 CTAs per SM)
mkdir -p gpurun_out
for lib in libf2d.so libf2d_gsmb5.so libf2d_gsmb6.so; do
  for cfg in "4096 20 3" "1024 80 3" "256 20 20"; do
    echo -n "$lib $cfg: " >> gpurun_out/gs_ab.log
    F2D_LIB_PATH=fluid-2d_b200/$lib timeout 120 python tools/gs_bench.py $cfg >> gpurun_out/gs_ab.log 2>&1
  done
done
cut -c1-260 gpurun_out/gs_ab.log
F2D_LIB_PATH=fluid-2d_b200/libf2d_gsmb6.so timeout 300 python -m pytest tests/test_gpu_cpu_semantics.py -x -q 2>&1 | tail -1
F2D_LIB_PATH=fluid-2d_b200/libf2d_gsmb5.so timeout 300 python -m pytest tests/test_gpu_cpu_semantics.py -x -q 2>&1 | tail -1
