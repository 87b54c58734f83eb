"""Per-phase cycle breakdown of the Gauss-Seidel wavefront (profiling build, -DF2D_GS_TIMING):
    make -C fluid-2d_b200/csrc OUT=$PWD/fluid-2d_b200/libf2d_gstime.so BUILD=$PWD/fluid-2d_b200/csrc/build_gstime VARIANT=-DF2D_GS_TIMING
    F2D_LIB_PATH=fluid-2d_b200/libf2d_gstime.so python tools/gs_timing.py N K"""
import ctypes as C
import json
import sys

sys.path.insert(0, ".")
import numpy as np  # noqa: E402

import fluid2d_b200 as f2d  # noqa: E402

n, k = int(sys.argv[1]), int(sys.argv[2])
L = f2d.load()
L.f2d_debug_gs_cycles.argtypes = [C.POINTER(C.c_uint64), C.c_int]
r = np.random.default_rng(0)
f = [r.standard_normal((n, n), dtype=np.float32) * np.float32(0.1) for _ in range(3)]
names = ["wait", "frame+commit", "prefetch_issue", "compute", "store", "release", "tiles"]
with f2d.FluidSolverB200.cpu_compatible(n, n, iters=k) as s:
    s.upload(*f)
    for diffuse in (False, True):
        out = (C.c_uint64 * 8)()
        s.bench_jacobi(diffuse, k, 1)
        L.f2d_debug_gs_cycles(out, 1)
        ms = s.bench_jacobi(diffuse, k, 1)  # warm-up launch + 1 timed launch
        L.f2d_debug_gs_cycles(out, 1)
        tiles = max(1, out[6])
        print(json.dumps({"n": n, "k": k, "diffuse": diffuse, "ms": ms,
                          "cycles_per_tile": {nm: round(out[i] / tiles, 1) for i, nm in enumerate(names[:6])}, "tiles": int(out[6])}))
