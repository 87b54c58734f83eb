#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n 2 gpurun_out/$name.log | cut -c1-300; }
TMO=600 run ncu_launches ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 3
TMO=600 run ncu_t8p ncu --set full --clock-control none --import-source on -k regex:k_jacobi_stream -s 12 -c 1 -f -o gpurun_out/jacobi_T8_pressure_final python tools/run_one.py 4096 80 8
TMO=600 run ncu_t8d ncu --set full --clock-control none --import-source on -k regex:k_jacobi_stream -s 12 -c 1 -f -o gpurun_out/jacobi_T8_diffuse_final python tools/run_one.py 4096 80 8 diffuse
TMO=600 run ncu_fused ncu --set full --clock-control none -k regex:k_jacobi_stream.*2, -s 3 -c 1 -f -o gpurun_out/jacobi_T8_fuseddiv_final python bench.py --steps 1 --warmup 3
TMO=600 run sanitizer compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_stages.py -q -m gpu -k "not full_size and (project or diffuse_exact) and (n8 or 8-)" -x
ls -la gpurun_out | head
