#!/bin/bash
# round 2, GPU call 7: two-operation constant division in the diffuse sweep (chaining removed again): tests, A/B, ncu, bench
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n 4 gpurun_out/$name.log | cut -c1-600; }
TMO=900 run tests_gpu python -m pytest tests -q -m gpu -x
TMO=600 run tests_div python -m pytest tests/test_gpu_stages.py -q -m gpu -x -s -k "corrected_divide"
TMO=600 run ab python tools/ab_variants.py 4096 80
TMO=600 run ncu_p ncu --set full --clock-control none --import-source on -k regex:k_jacobi_stream -s 12 -c 1 -f -o gpurun_out/jacobi_T8_pressure_v4 python tools/run_one.py 4096 80 8
TMO=600 run ncu_d ncu --set full --clock-control none --import-source on -k regex:k_jacobi_stream -s 12 -c 1 -f -o gpurun_out/jacobi_T8_diffuse_v4 python tools/run_one.py 4096 80 8 diffuse
TMO=900 run bench_1gpu python bench.py --steps 20 --warmup 5
