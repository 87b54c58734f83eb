#!/bin/bash
# GPU run: pipelined solve() parity + e2e A/B (F2D_HOST_PIPELINE=1 vs 0) through bench.py
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_step.py -x -q -k "pipelined or solve_host or graph_replay" > gpurun_out/test_pipe.log 2>&1
echo "pytest rc=$?" >> gpurun_out/test_pipe.log
tail -4 gpurun_out/test_pipe.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_pipe1.log 2>&1
F2D_HOST_PIPELINE=0 timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_pipe0.log 2>&1
python - <<'PY'
import json
for f in ("gpurun_out/bench_pipe1.log", "gpurun_out/bench_pipe0.log"):
    try:
        l = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step", round(l["ms_per_step"], 3), "e2e ms", round(l["e2e"]["ms_per_step"], 3), "e2e value", l["e2e"]["value"])
    except Exception as e:
        print(f, "unreadable:", e, open(f).read()[-400:])
PY
