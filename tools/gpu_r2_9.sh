#!/bin/bash
# round 2, GPU call 9: out_prev removed (bottom chunk owns the row above the edge), rhs-generation code dropped; m0 A/B
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n 4 gpurun_out/$name.log | cut -c1-600; }
TMO=900 run tests_gpu python -m pytest tests -q -m gpu -x
TMO=600 run ab python tools/ab_variants.py 4096 80
TMO=900 run bench_1gpu python bench.py --steps 20 --warmup 5
