#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gs_relax -s 1 -c 1 -f -o gpurun_out/gs_relax_r01 python tools/gs_bench.py 1024 20 1 > gpurun_out/ncu_gs.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_gs.log; ls -la gpurun_out/gs_relax_r01.ncu-rep
