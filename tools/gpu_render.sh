#!/bin/bash
mkdir -p gpurun_out/golden
timeout 200 python tests/golden/make_refrender_fixtures.py gpurun_out/golden 2>&1 | tail -3
timeout 200 python -m pytest tests/test_gpu_render_ref.py -x -q 2>&1 | tail -3
ls -la gpurun_out/golden | tail -3
