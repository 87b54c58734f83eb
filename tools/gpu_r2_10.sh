#!/bin/bash
# round 2, GPU call 10: plain rhs ring as default; programmatic dependent launch of the relaxation passes (F2D_STREAM_PDL) A/B
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n 4 gpurun_out/$name.log | cut -c1-600; }
TMO=900 run tests_gpu python -m pytest tests -q -m gpu -x
F2D_STREAM_PDL=1 TMO=900 run tests_gpu_pdl python -m pytest tests -q -m gpu -x
F2D_STREAM_PDL=0 TMO=600 run ab_pdl0 python tools/ab_variants.py 4096 80
F2D_STREAM_PDL=1 TMO=600 run ab_pdl1 python tools/ab_variants.py 4096 80
F2D_STREAM_PDL=0 TMO=300 run small_pdl0 python tools/tune_small.py
F2D_STREAM_PDL=1 TMO=300 run small_pdl1 python tools/tune_small.py
