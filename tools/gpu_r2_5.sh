#!/bin/bash
# round 2, GPU call 5: chained passes (F2D_STREAM_CHAIN) and scaled pressure levels (ps0 = unscaled build): tests, A/B, bench
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n 4 gpurun_out/$name.log | cut -c1-600; }
TMO=900 run tests_gpu python -m pytest tests -q -m gpu -x
F2D_STREAM_CHAIN=0 TMO=600 run ab_chain0 python tools/ab_variants.py 4096 80
F2D_STREAM_CHAIN=1 TMO=600 run ab_chain1 python tools/ab_variants.py 4096 80
TMO=300 run small python tools/tune_small.py
TMO=600 run racecheck_gs compute-sanitizer --tool racecheck --kernel-name kns=k_gs_relax python -m pytest tests/test_gpu_cpu_semantics.py -q -m gpu -x -k "gauss_seidel_diffuse and not many and not non_square"
TMO=600 run synccheck_gs compute-sanitizer --tool synccheck --kernel-name kns=k_gs_relax python -m pytest tests/test_gpu_cpu_semantics.py -q -m gpu -x -k "gauss_seidel_diffuse and not many and not non_square"
TMO=600 run synccheck_stream compute-sanitizer --tool synccheck --kernel-name kns=k_jacobi_stream python -m pytest tests/test_gpu_stages.py -q -m gpu -x -k "diffuse or project"
TMO=900 run bench_1gpu python bench.py --steps 20 --warmup 5
