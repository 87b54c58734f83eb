#!/bin/bash
# round 2, GPU call 12: float4 advect / scatter kernels: tests, ncu --set full of both, launch list, bench
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n 4 gpurun_out/$name.log | cut -c1-600; }
TMO=900 run tests_gpu python -m pytest tests -q -m gpu -x
F2D_BENCH_SCALING_BASE=0 TMO=600 run ncu_adv ncu --set full --clock-control none --import-source on -k regex:"k_advect_velocity|k_scatter_density" -s 6 -c 2 -f -o gpurun_out/advect_scatter_v4 python bench.py --steps 2 --warmup 3
F2D_BENCH_SCALING_BASE=0 TMO=600 run ncu_launches ncu --metrics gpu__time_duration.sum --clock-control none -s 105 -c 80 --csv --log-file gpurun_out/launches_r02_v6.csv python bench.py --steps 2 --warmup 3
TMO=900 run bench_1gpu python bench.py --steps 20 --warmup 5
