#!/bin/bash
mkdir -p gpurun_out
for cfg in "256 20" "1024 20" "4096 20"; do
  F2D_LIB_PATH=fluid-2d_b200/libf2d_gstime.so timeout 120 python tools/gs_timing.py $cfg >> gpurun_out/gs_timing.log 2>&1
done
cat gpurun_out/gs_timing.log
