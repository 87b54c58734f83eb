#!/bin/bash
# round 2, GPU call 14: dependent launch for every kernel of the step: tests, small-grid sweep with PDL on/off, 4096^2 A/B
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n 3 gpurun_out/$name.log | cut -c1-400; }
TMO=900 run tests_gpu python -m pytest tests -q -m gpu -x
F2D_STREAM_PDL=0 TMO=300 run small2_pdl0 python tools/tune_small.py
F2D_STREAM_PDL=1 TMO=300 run small2_pdl1 python tools/tune_small.py
F2D_STREAM_PDL=0 TMO=300 run ab2_pdl0 python tools/ab_variants.py 4096 80
F2D_STREAM_PDL=1 TMO=300 run ab2_pdl1 python tools/ab_variants.py 4096 80
