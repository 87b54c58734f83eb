#!/bin/bash
# GPU session 2 (1 GPU): all gpu tests, bench, tuning sweep at 4096^2 and 16384^2, ncu of the T=8 and T=4 kernels.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n 8 gpurun_out/$name.log; }
TMO=300 run smoke python -c "import __graft_entry__ as g; g.smoke()"
TMO=1200 run tests python -m pytest tests -q -m gpu --maxfail=20
TMO=600 run bench python bench.py --steps 10 --warmup 3
TMO=600 run bench_t4 env F2D_TEMPORAL_BLOCK=4 python bench.py --steps 10 --warmup 3
TMO=900 run tune4096 python tools/tune_stream.py 4096 80
TMO=900 run tune16384 python tools/tune_stream.py 16384 80
TMO=300 run headless bash -c "g++ -std=c++14 -O2 -Iinclude examples/headless_sim.cpp -Lfluid-2d_b200 -lf2d -Wl,-rpath,\$PWD/fluid-2d_b200 -o /tmp/headless_sim && /tmp/headless_sim 256 100 && /tmp/headless_sim 1024 20"
TMO=600 run ncu_launches ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r01b.csv python bench.py --steps 1 --warmup 3
TMO=600 run ncu_t8 ncu --set full --clock-control none --import-source on -k regex:k_jacobi_stream -s 12 -c 2 -f -o gpurun_out/jacobi_T8_r01b python tools/run_one.py 4096 80 8
TMO=600 run ncu_t4 ncu --set full --clock-control none --import-source on -k regex:k_jacobi_stream -s 22 -c 2 -f -o gpurun_out/jacobi_T4_r01b python tools/run_one.py 4096 80 4
TMO=600 run ncu_t4d ncu --set full --clock-control none --import-source on -k regex:k_jacobi_stream -s 22 -c 1 -f -o gpurun_out/jacobi_T4_diffuse_r01b python tools/run_one.py 4096 80 4 diffuse
ls -la gpurun_out
