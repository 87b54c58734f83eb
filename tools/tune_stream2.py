"""Planner sweep of the streaming Jacobi kernel at T=8 (run on the GPU box): edge-strip cost, warps per CTA.
python tools/tune_stream2.py 4096 80"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import fluid2d_b200 as f2d  # noqa: E402

KEYS = ("F2D_STREAM_CHUNK_ROWS", "F2D_STREAM_WARPS_PER_CTA", "F2D_STREAM_EDGE_COST_PCT", "F2D_STREAM_MIN_CHUNK_MULT", "F2D_STREAM_PDL")


def run(n, k, f, env):
    for key in KEYS:
        os.environ.pop(key, None)
    os.environ.update({a: str(b) for a, b in env.items()})
    with f2d.FluidSolverB200(n, n, diffuse_iters=k, project_iters=k, temporal_block=8, temporal_block_diffuse=8) as s:
        s.upload(*f)
        p = min(s.bench_jacobi(False, k, 3) for _ in range(3)) / 3
        d = min(s.bench_jacobi(True, k, 3) for _ in range(3)) / 3
        s.step(0.5, 1e-6, 0.02, 2)
        st = min(s.step_timed(0.5, 1e-6, 0.02, 3) for _ in range(3)) / 3
    print(json.dumps(dict(n=n, k=k, env=env, pressure_ms=round(p, 4), diffuse_ms=round(d, 4), step_ms=round(st, 3))), flush=True)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 80
    r = np.random.default_rng(0)
    f = [r.standard_normal((n, n), dtype=np.float32) * np.float32(0.1) for _ in range(3)]
    run(n, k, f, {})
    for pct in (100, 115, 125, 150, 175, 200):
        run(n, k, f, {"F2D_STREAM_EDGE_COST_PCT": pct})
    for w in (1, 2, 3):
        run(n, k, f, {"F2D_STREAM_WARPS_PER_CTA": w})
    for rows in (64, 72, 80, 96, 112, 128):
        run(n, k, f, {"F2D_STREAM_CHUNK_ROWS": rows})
    run(n, k, f, {"F2D_STREAM_PDL": 0})


if __name__ == "__main__":
    main()
