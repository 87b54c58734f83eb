#!/bin/bash
# round 2, GPU call 11: final streaming kernel: ncu --set full of both passes, launch list of the bench step, parity report
# incl. 4096^2 and 16384^2 against the live reference GPU solver, bench
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n 4 gpurun_out/$name.log | cut -c1-600; }
TMO=600 run ncu_p ncu --set full --clock-control none --import-source on -k regex:k_jacobi_stream -s 12 -c 1 -f -o gpurun_out/jacobi_T8_pressure_v5 python tools/run_one.py 4096 80 8
TMO=600 run ncu_d ncu --set full --clock-control none --import-source on -k regex:k_jacobi_stream -s 12 -c 1 -f -o gpurun_out/jacobi_T8_diffuse_v5 python tools/run_one.py 4096 80 8 diffuse
F2D_BENCH_SCALING_BASE=0 TMO=600 run ncu_launches ncu --metrics gpu__time_duration.sum --clock-control none -s 105 -c 80 --csv --log-file gpurun_out/launches_r02_v5.csv python bench.py --steps 2 --warmup 3
TMO=1500 run parity python tools/parity_report.py gpurun_out/parity_report_r02.md --big
TMO=900 run bench_1gpu python bench.py --steps 20 --warmup 5
