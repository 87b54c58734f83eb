"""Run one relaxation configuration a few times (for ncu): python tools/run_one.py N K T [diffuse]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import fluid2d_b200 as f2d  # noqa: E402

n, k, T = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
diffuse = len(sys.argv) > 4 and sys.argv[4] == "diffuse"
r = np.random.default_rng(0)
f = [r.standard_normal((n, n), dtype=np.float32) * np.float32(0.1) for _ in range(3)]
with f2d.FluidSolverB200(n, n, diffuse_iters=k, project_iters=k, temporal_block=T, temporal_block_diffuse=T) as s:
    s.upload(*f)
    ms = s.bench_jacobi(diffuse, k, 3) / 3
    print("n=%d K=%d T=%d %s: %.4f ms per solve, %.1f GB/s algorithmic" % (n, k, T, "diffuse" if diffuse else "pressure", ms, 12.0 * n * n * k / ms / 1e6))
