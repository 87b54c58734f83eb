"""A/B of f2d_solve_host: pipelined (F2D_HOST_PIPELINE=1) vs serial (=0), by grid size.  One subprocess per mode
(the switch is read at solver creation).  python tools/solve_ab.py"""
import json
import os
import subprocess
import sys

CHILD = r'''
import sys, time, json
sys.path.insert(0, ".")
import numpy as np, torch
import fluid2d_b200 as f2d
out = {}
for n, kd, kp in ((256, 15, 20), (512, 15, 20), (1024, 40, 40), (2048, 40, 40), (4096, 80, 80)):
    r = np.random.default_rng(n)
    hs = [torch.from_numpy((r.standard_normal((n, n)) * 0.01).astype(np.float32)).pin_memory().numpy() for _ in range(6)]
    with f2d.FluidSolverB200(n, n, diffuse_iters=kd, project_iters=kp) as s:
        for _ in range(3):
            s.solve(hs[0], hs[3], 0.5, hs[1], hs[2], hs[4], hs[5], 1e-6, 0.02)
        reps = 20 if n <= 1024 else 8
        t = time.perf_counter()
        for _ in range(reps):
            s.solve(hs[0], hs[3], 0.5, hs[1], hs[2], hs[4], hs[5], 1e-6, 0.02)
        out[str(n)] = round(1e3 * (time.perf_counter() - t) / reps, 4)
print(json.dumps(out))
'''
for mode in ("1", "0"):
    env = dict(os.environ, F2D_HOST_PIPELINE=mode, F2D_HOST_PIPELINE_MIN_BYTES="0")
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    print("pipeline=%s" % mode, r.stdout.strip() or r.stderr[-600:], flush=True)
