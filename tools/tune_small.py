"""Small-grid sweep: step time at 256^2 (K=20) and 1024^2 (K=40) for T and the minimum chunk height."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import fluid2d_b200 as f2d
for n, k in ((256, 20), (512, 20), (1024, 40), (2048, 40)):
    r = np.random.default_rng(0)
    f = [r.standard_normal((n, n), dtype=np.float32) * np.float32(0.1) for _ in range(3)]
    for T in (2, 4, 8):
        for mult in (1, 2, 4):
            for graph in (True,):
                os.environ["F2D_STREAM_MIN_CHUNK_MULT"] = str(mult)
                with f2d.FluidSolverB200(n, n, diffuse_iters=k, project_iters=k, temporal_block=T, temporal_block_diffuse=T, use_graph=graph) as s:
                    s.upload(*f)
                    s.step(0.5, 1e-6, 0.02, 5)
                    ms = min(s.step_timed(0.5, 1e-6, 0.02, 20) for _ in range(3)) / 20
                print(json.dumps(dict(n=n, k=k, T=T, min_chunk_mult=mult, step_ms=round(ms, 4), Gcs=round(n * n / ms / 1e6, 3))), flush=True)
