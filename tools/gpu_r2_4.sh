#!/bin/bash
# round 2, GPU call 4: A/B of the rhs-generation streaming kernel (g0m1 = v3 baseline, default = g1m1, g1m0),
# the full -m gpu suite on the new default, racecheck / synccheck on k_gs_relax, ncu of both passes, bench
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n 4 gpurun_out/$name.log | cut -c1-600; }
TMO=600 run ab_r2_4 python tools/ab_variants.py 4096 80
TMO=900 run tests_gpu python -m pytest tests -q -m gpu -x
F2D_LIB_PATH=$PWD/fluid-2d_b200/libf2d_g1m0.so TMO=600 run tests_g1m0 python -m pytest tests/test_gpu_stages.py tests/test_gpu_step.py -q -m gpu -x
TMO=600 run ncu_p ncu --set full --clock-control none --import-source on -k regex:k_jacobi_stream -s 12 -c 1 -f -o gpurun_out/jacobi_T8_pressure_g1m1 python tools/run_one.py 4096 80 8
TMO=600 run ncu_d ncu --set full --clock-control none --import-source on -k regex:k_jacobi_stream -s 12 -c 1 -f -o gpurun_out/jacobi_T8_diffuse_g1m1 python tools/run_one.py 4096 80 8 diffuse
TMO=600 run racecheck_gs compute-sanitizer --tool racecheck --kernel-name regex:k_gs_relax python -m pytest tests/test_gpu_cpu_semantics.py -q -m gpu -x -k "gauss_seidel_diffuse and not many and not non_square"
TMO=600 run synccheck_gs compute-sanitizer --tool synccheck --kernel-name regex:k_gs_relax python -m pytest tests/test_gpu_cpu_semantics.py -q -m gpu -x -k "gauss_seidel_diffuse and not many and not non_square"
TMO=900 run bench_1gpu python bench.py --steps 20 --warmup 5
ls -la gpurun_out | head -30
