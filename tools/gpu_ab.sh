#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== tests (default lib)"; timeout 1200 python -m pytest tests -q -m gpu --maxfail=20 > gpurun_out/tests.log 2>&1; echo "exit $?"; tail -n 3 gpurun_out/tests.log
echo "=== ab 4096"; timeout 900 python tools/ab_variants.py 4096 80 > gpurun_out/ab_4096.log 2>&1; cat gpurun_out/ab_4096.log
echo "=== ab 16384"; timeout 900 python tools/ab_variants.py 16384 80 > gpurun_out/ab_16384.log 2>&1; cat gpurun_out/ab_16384.log
