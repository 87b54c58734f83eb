#!/bin/bash
# GPU session (1 GPU): validate planner / shuffle pipelining / vector kernels / e2e overlap; bench; profiles.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n 6 gpurun_out/$name.log; }
TMO=300 run smoke python -c "import __graft_entry__ as g; g.smoke()"
TMO=1200 run tests python -m pytest tests -q -m gpu --maxfail=20
TMO=600 run bench python bench.py --steps 10 --warmup 3
TMO=900 run tune4096 python tools/tune_stream.py 4096 80
TMO=600 run bench16384 python bench.py --size 16384 --steps 5 --warmup 3
TMO=600 run bench1024 python bench.py --size 1024 --iters 40 --steps 50 --warmup 5
TMO=600 run bench256 python bench.py --size 256 --iters 20 --steps 100 --warmup 5
TMO=600 run ncu_launches ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r01c.csv python bench.py --steps 1 --warmup 3
TMO=600 run ncu_t8 ncu --set full --clock-control none --import-source on -k regex:k_jacobi_stream -s 12 -c 1 -f -o gpurun_out/jacobi_T8_r01c python tools/run_one.py 4096 80 8
TMO=600 run ncu_t4d ncu --set full --clock-control none --import-source on -k regex:k_jacobi_stream -s 22 -c 1 -f -o gpurun_out/jacobi_T4_diffuse_r01c python tools/run_one.py 4096 80 4 diffuse
ls -la gpurun_out | head -30
