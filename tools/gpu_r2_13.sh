#!/bin/bash
# round 2, GPU call 13: level split (two independent chains per row step) vs ls0; tests
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n 4 gpurun_out/$name.log | cut -c1-600; }
TMO=900 run tests_gpu python -m pytest tests -q -m gpu -x
TMO=600 run ab python tools/ab_variants.py 4096 80
