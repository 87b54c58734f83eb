#!/bin/bash
# round 2, GPU call 2: packed-fp32x2 streaming kernel -- parity (stage tests, 4096^2 property, headline config) and timing
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n 3 gpurun_out/$name.log | cut -c1-400; }
TMO=900 run tests_stages python -m pytest tests/test_gpu_stages.py tests/test_gpu_step.py -q -m gpu -x
for ec in 100 135 170; do
  echo "--- edge cost $ec"
  F2D_STREAM_EDGE_COST_PCT=$ec python tools/run_one.py 4096 80 8
  F2D_STREAM_EDGE_COST_PCT=$ec python tools/run_one.py 4096 80 8 diffuse
done
python tools/run_one.py 4096 80 4
python tools/run_one.py 4096 80 4 diffuse
python tools/run_one.py 1024 40 4
python tools/run_one.py 256 20 4
F2D_BENCH_QUICK=1 F2D_BENCH_SCALING_BASE=0 TMO=600 run bench_quick python bench.py --steps 20 --warmup 5
