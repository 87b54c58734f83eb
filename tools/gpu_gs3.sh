#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cpu_semantics.py tests/test_headless_app.py -x -q -m gpu > gpurun_out/test_cpu_sem.log 2>&1
echo "pytest rc=$?" >> gpurun_out/test_cpu_sem.log; tail -3 gpurun_out/test_cpu_sem.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cpu_sem_4096_K20.csv python tools/gs_bench.py 4096 20 1 > gpurun_out/ncu_gs_launches.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/launches_cpu_sem_4096_K20.csv
