#!/bin/bash
mkdir -p gpurun_out
SECONDS=0
timeout 600 python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err
echo "rc=$? wall=${SECONDS}s"; tail -n 1 gpurun_out/bench_default.log | cut -c1-200; tail -3 gpurun_out/bench_default.err
python - <<'PY'
import json
try:
    l = json.loads(open("gpurun_out/bench_default.log").read().strip().splitlines()[-1])
    print(l["scaling_base"]); print(l["e2e"]["ms_per_step"], l["cpu_exact_mode"]["ms_per_step"])
except Exception as e:
    print("unreadable", e)
PY
