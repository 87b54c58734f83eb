#!/bin/bash
# round 2, final single-GPU sequence: the full -m gpu suite, smoke, both bench arms
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n 3 gpurun_out/$name.log | cut -c1-400; }
TMO=900 run tests_gpu_final python -m pytest tests -q -m gpu -x
TMO=300 run smoke_final python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
TMO=900 run bench_ref_final python bench.py --impl reference --steps 3 --warmup 1
TMO=900 run bench_final python bench.py --steps 20 --warmup 5
