#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/solve_ab.py > gpurun_out/solve_ab.log 2>&1; cat gpurun_out/solve_ab.log
timeout 300 python -m pytest tests/test_gpu_step.py -x -q -k "pipelined or solve_host" 2>&1 | tail -2
