"""A/B: time the pressure / diffuse solves and a full step for each built libf2d variant (subprocess per
variant so each gets its own library).  python tools/ab_variants.py 4096 80"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json
sys.path.insert(0, %r)
import numpy as np
import fluid2d_b200 as f2d
n, k = int(sys.argv[1]), int(sys.argv[2])
r = np.random.default_rng(0)
f = [r.standard_normal((n, n), dtype=np.float32) * np.float32(0.1) for _ in range(3)]
out = {}
for T in (8, 4):
    with f2d.FluidSolverB200(n, n, diffuse_iters=k, project_iters=k, temporal_block=T, temporal_block_diffuse=T) as s:
        s.upload(*f)
        p = min(s.bench_jacobi(False, k, 3) for _ in range(2)) / 3
        d = min(s.bench_jacobi(True, k, 3) for _ in range(2)) / 3
        s.step(0.5, 1e-6, 0.02, 2)
        st = min(s.step_timed(0.5, 1e-6, 0.02, 3) for _ in range(2)) / 3
        out["T%%d" %% T] = dict(pressure_ms=round(p, 4), diffuse_ms=round(d, 4), step_ms=round(st, 3))
print(json.dumps(out))
''' % ROOT


def main():
    n, k = sys.argv[1], sys.argv[2]
    libdir = os.path.join(ROOT, "fluid-2d_b200")
    for name in sorted(x for x in os.listdir(libdir) if x.startswith("libf2d") and x.endswith(".so")):
        env = dict(os.environ, F2D_LIB_PATH=os.path.join(libdir, name))
        r = subprocess.run([sys.executable, "-c", CHILD, n, k], env=env, capture_output=True, text=True)
        print(name, r.stdout.strip() or r.stderr[-500:], flush=True)


if __name__ == "__main__":
    main()
