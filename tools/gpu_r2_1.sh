#!/bin/bash
# round 2, GPU call 1: micro-rates for the packed-fp32 decision, the new parity tests, the full -m gpu suite, bench with the ref_gpu leg
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n 3 gpurun_out/$name.log | cut -c1-400; }
TMO=120 run ubench_fp32x2 tools/ubench/fp32x2_rates
TMO=900 run tests_new python -m pytest tests/test_gpu_step.py tests/test_gpu_stages.py -q -m gpu -x -s -k "headline or published"
TMO=900 run bench_1gpu python bench.py --steps 20 --warmup 5
TMO=1500 run tests_gpu python -m pytest tests -q -m gpu -x
