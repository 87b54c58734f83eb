"""Times the fluid_solver_cpu-compatible mode (F2D_SEM_CPU): whole steps and the Gauss-Seidel wavefront alone.
usage: python tools/gs_bench.py N K [steps]   -> one JSON line"""
import json
import sys

sys.path.insert(0, ".")
import fluid2d_b200 as f2d  # noqa: E402
from oracle import sfo  # noqa: E402  (canonical input fields only)

n, k = int(sys.argv[1]), int(sys.argv[2])
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
d, u, v, sd, su, sv = sfo.canonical_fields(n)
with f2d.FluidSolverB200.cpu_compatible(n, n, iters=k) as s:
    s.upload(d, u, v)
    s.set_sources(sd, su, sv)
    s.step(0.5, 1e-6, 0.02, 2)
    s.sync()
    ms = s.step_timed(0.5, 1e-6, 0.02, steps) / steps
    s.sync()
    gs_p = s.bench_jacobi(False, k, 3) / 3
    gs_d = s.bench_jacobi(True, k, 3) / 3
    s.sync()
print(json.dumps({"n": n, "k": k, "ms_per_step": ms, "cell_steps_per_s": n * n / (ms * 1e-3),
                  "gs_pressure_ms": gs_p, "gs_diffuse_ms": gs_d,
                  "gs_pressure_gcell_sweeps_per_s": n * n * k / (gs_p * 1e-3) / 1e9}))
