"""Parameter sweep of the streaming Jacobi kernel + comparison points (run on the GPU box).
Prints one line per configuration: ms per pressure / diffuse solve of K sweeps, algorithmic GB/s."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import fluid2d_b200 as f2d  # noqa: E402


def fields(n):
    r = np.random.default_rng(0)
    return [r.standard_normal((n, n), dtype=np.float32) * np.float32(0.1) for _ in range(3)]


def run(n, k, reps, mode, T, div, env):
    for key in ("F2D_STREAM_CHUNK_ROWS", "F2D_STREAM_WARPS_PER_CTA", "F2D_STREAM_RHS_SMEM", "F2D_STREAM_MIN_BLOCKS"):
        os.environ.pop(key, None)
    os.environ.update({a: str(b) for a, b in env.items()})
    d, u, v = FIELDS[n]
    with f2d.FluidSolverB200(n, n, diffuse_iters=k, project_iters=k, jacobi_mode=mode, temporal_block=T, temporal_block_diffuse=T,
                             divide_mode=div) as s:
        s.upload(d, u, v)
        p_ms = s.bench_jacobi(False, k, reps) / reps
        d_ms = s.bench_jacobi(True, k, reps) / reps
        s.step(0.5, 1e-6, 0.02, 2)
        st_ms = s.step_timed(0.5, 1e-6, 0.02, 3) / 3
    gb = 12.0 * n * n * k / 1e9
    rec = dict(n=n, k=k, mode=mode, T=T, div=div, env=env, pressure_ms=round(p_ms, 4), diffuse_ms=round(d_ms, 4),
               pressure_GBs=round(gb / (p_ms * 1e-3), 1), diffuse_GBs=round(gb / (d_ms * 1e-3), 1),
               step_ms=round(st_ms, 3), cell_steps_per_s=round(n * n / (st_ms * 1e-3) / 1e9, 4))
    print(json.dumps(rec), flush=True)


FIELDS = {}


def main():
    sizes = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "4096").split(",")]
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 80
    for n in sizes:
        FIELDS[n] = fields(n)
        reps = 3 if n <= 8192 else 1
        for T in (1, 2, 4, 8):
            run(n, k, reps, 1, T, 1, {})
        run(n, k, reps, 1, 8, 1, {"F2D_STREAM_MIN_BLOCKS": 2})
        run(n, k, reps, 1, 4, 1, {"F2D_STREAM_RHS_SMEM": 1})
        run(n, k, reps, 1, 4, 1, {"F2D_STREAM_RHS_SMEM": 1, "F2D_STREAM_MIN_BLOCKS": 5})
        run(n, k, reps, 1, 4, 1, {"F2D_STREAM_WARPS_PER_CTA": 2})
        run(n, k, reps, 1, 8, 1, {"F2D_STREAM_WARPS_PER_CTA": 2})
        FIELDS.pop(n)
    # the unmodified reference GPU solver on the same box ("the kernel to beat"), its own solve()
    try:
        from oracle import refs
        if refs.have_gpu():
            g = refs.ref_gpu()
            for n in ():
                f = [np.zeros((n, n), np.float32) + np.float32(0.1) for _ in range(6)]
                g.solve(f[0], f[3], 0.5, f[1], f[2], f[4], f[5], 1e-6, 0.02, 1)
                t = time.perf_counter()
                *_, ms = g.solve(f[0], f[3], 0.5, f[1], f[2], f[4], f[5], 1e-6, 0.02, 3)
                wall = (time.perf_counter() - t) / 3
                print(json.dumps(dict(ref_gpu_solve=True, n=n, kd=15, kp=20, ms_per_step=round(ms / 3, 3),
                                      wall_ms_per_step=round(wall * 1e3, 3),
                                      cell_steps_per_s=round(n * n / (ms / 3 * 1e-3) / 1e9, 5))), flush=True)
    except Exception as e:  # noqa: BLE001
        print("ref_gpu unavailable:", e)


if __name__ == "__main__":
    main()
