#!/bin/bash
# round 2, GPU call 8: out_prev removed; pressure rhs generations (pg6 / pg3) A/B; tests
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n 4 gpurun_out/$name.log | cut -c1-600; }
TMO=900 run tests_gpu python -m pytest tests -q -m gpu -x
F2D_LIB_PATH=$PWD/fluid-2d_b200/libf2d_pg6.so TMO=600 run tests_pg6 python -m pytest tests/test_gpu_stages.py -q -m gpu -x -k project
TMO=600 run ab python tools/ab_variants.py 4096 80
