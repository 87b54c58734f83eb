#!/bin/bash
# First GPU session: smoke, fixtures, parity tests, bench, tuning sweep, ncu.  Run under gpurun.
set -u
mkdir -p gpurun_out/golden
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n 6 gpurun_out/$name.log; }
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
TMO=300 run smoke python -c "import __graft_entry__ as g; g.smoke()"
TMO=300 run fixtures python tests/golden/make_refgpu_fixtures.py gpurun_out/golden
TMO=900 run test_stages python -m pytest tests/test_gpu_stages.py -q -m gpu --maxfail=20 -k "not full_size"
TMO=600 run test_step python -m pytest tests/test_gpu_step.py -q -m gpu --maxfail=20
TMO=600 run test_fullsize python -m pytest tests/test_gpu_stages.py -q -m gpu -k "full_size"
TMO=600 run bench python bench.py --steps 10 --warmup 3
TMO=900 run tune python tools/tune_stream.py 4096 80
TMO=600 run ncu_launches ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 1 --warmup 3
TMO=600 run ncu_full ncu --set full --clock-control none --import-source on -k regex:k_jacobi_stream -s 20 -c 3 -f -o gpurun_out/jacobi_stream_r01 python bench.py --steps 1 --warmup 3
ls -la gpurun_out
