#!/bin/bash
# round 2, GPU call 6: chained passes A/B after the zero-start fix, racecheck on full vs partial bands, bench
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n 4 gpurun_out/$name.log | cut -c1-600; }
TMO=900 run tests_gpu python -m pytest tests -q -m gpu -x
F2D_STREAM_CHAIN=0 TMO=600 run ab_chain0 python tools/ab_variants.py 4096 80
F2D_STREAM_CHAIN=1 TMO=600 run ab_chain1 python tools/ab_variants.py 4096 80
F2D_STREAM_CHAIN=0 TMO=300 run small_chain0 python tools/tune_small.py
F2D_STREAM_CHAIN=1 TMO=300 run small_chain1 python tools/tune_small.py
TMO=300 run racecheck_gs_n34 compute-sanitizer --tool racecheck --kernel-name kns=k_gs_relax python -m pytest tests/test_gpu_cpu_semantics.py -q -m gpu -x -k "test_gauss_seidel_diffuse and 34"
TMO=300 run racecheck_gs_n33 compute-sanitizer --tool racecheck --kernel-name kns=k_gs_relax python -m pytest tests/test_gpu_cpu_semantics.py -q -m gpu -x -k "test_gauss_seidel_diffuse and 33"
TMO=900 run bench_1gpu python bench.py --steps 20 --warmup 5
