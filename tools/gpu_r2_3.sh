#!/bin/bash
# round 2, GPU call 3: ncu of the packed-fp32x2 pressure / diffuse passes
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n 3 gpurun_out/$name.log | cut -c1-400; }
TMO=600 run ncu_p ncu --set full --clock-control none --import-source on -k regex:k_jacobi_stream -s 12 -c 1 -f -o gpurun_out/jacobi_T8_pressure_v3 python tools/run_one.py 4096 80 8
TMO=600 run ncu_d ncu --set full --clock-control none --import-source on -k regex:k_jacobi_stream -s 12 -c 1 -f -o gpurun_out/jacobi_T8_diffuse_v3 python tools/run_one.py 4096 80 8 diffuse
ls -la gpurun_out | head -20
