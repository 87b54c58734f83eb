#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n 4 gpurun_out/$name.log | cut -c1-400; }
TMO=300 run smoke python -c "import __graft_entry__ as g; g.smoke()"
TMO=1200 run tests python -m pytest tests -q -m gpu --maxfail=20
TMO=600 run bench python bench.py --steps 10 --warmup 3
TMO=900 run tune4096 python tools/tune_stream.py 4096 80
TMO=900 run tune16384 python tools/tune_stream.py 16384 80
ls gpurun_out | head -30
