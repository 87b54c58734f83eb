"""ctypes binding of the plain-C oracle (oracle/stable_fluids_oracle.c).  TEST INFRASTRUCTURE ONLY.

All functions take/return C-contiguous float32 numpy arrays of shape (rows, cols) and work on
copies unless stated otherwise.  `sem` selects the arithmetic restated: SEM_GPU follows
src/fluid_solver_gpu.cu (the parity target), SEM_CPU follows src/fluid_solver_cpu.cpp.
"""
import ctypes as C
import os
import subprocess
import threading

import numpy as np

SEM_GPU = 0
SEM_CPU = 1
BND_CONTINUOUS = 0
BND_OPPOSITE_HORIZONTAL = 1
BND_OPPOSITE_VERTICAL = 2

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None
_FP = C.POINTER(C.c_float)


def build(force=False):
    """Compile liboracle.so (gcc) and, when /root/reference is present, oracle/_ref."""
    src = os.path.join(_HERE, "stable_fluids_oracle.c")
    stale = (not os.path.exists(_LIB_PATH)) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)
    if force or stale or os.path.isdir("/root/reference/src"):
        subprocess.run(["make", "-C", _HERE, "all"], check=True, stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        sz, f, i, u = C.c_size_t, C.c_float, C.c_int, C.c_uint
        L.sfo_set_bnd.argtypes = [_FP, sz, sz, i, i]
        L.sfo_add_sources.argtypes = [_FP, _FP, sz, sz, f, i]
        L.sfo_diffuse_coeff.argtypes = [sz, sz, f, f]
        L.sfo_diffuse_coeff.restype = f
        L.sfo_dt0.argtypes = [sz, sz, f, i]
        L.sfo_dt0.restype = f
        L.sfo_diffuse.argtypes = [_FP, sz, sz, i, f, f, u, i]
        L.sfo_advect_gather.argtypes = [_FP, _FP, _FP, sz, sz, i, f, i]
        L.sfo_advect_scatter.argtypes = [_FP, _FP, _FP, sz, sz, i, f, i]
        L.sfo_smooth.argtypes = [_FP, sz, sz]
        L.sfo_project.argtypes = [_FP, _FP, sz, sz, u, i, _FP, _FP]
        L.sfo_steps.argtypes = [_FP, _FP, f, _FP, _FP, _FP, _FP, f, f, sz, sz, u, u, i, i, u]
        L.sfo_canonical_fields.argtypes = [sz, sz, sz, _FP, _FP, _FP, _FP, _FP, _FP]
        L.sfo_fnv1a64.argtypes = [C.c_void_p, sz]
        L.sfo_fnv1a64.restype = C.c_uint64
        for name in ("sfo_set_bnd", "sfo_add_sources", "sfo_diffuse", "sfo_advect_gather",
                     "sfo_advect_scatter", "sfo_smooth", "sfo_project", "sfo_steps",
                     "sfo_canonical_fields"):
            getattr(L, name).restype = None
        _lib = L
    return _lib


def _p(a):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_FP)


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float32).copy()


def set_bnd(f, kind, sem=SEM_GPU):
    f = _c(f)
    lib().sfo_set_bnd(_p(f), f.shape[0], f.shape[1], kind, sem)
    return f


def add_sources(f, s, dt, sem=SEM_GPU):
    f, s = _c(f), _c(s)
    lib().sfo_add_sources(_p(f), _p(s), f.shape[0], f.shape[1], dt, sem)
    return f


def diffuse(f, kind, rate, dt, iters, sem=SEM_GPU):
    f = _c(f)
    lib().sfo_diffuse(_p(f), f.shape[0], f.shape[1], kind, rate, dt, iters, sem)
    return f


def advect_gather(f, u, v, kind, dt, sem=SEM_GPU):
    f, u, v = _c(f), _c(u), _c(v)
    lib().sfo_advect_gather(_p(f), _p(u), _p(v), f.shape[0], f.shape[1], kind, dt, sem)
    return f


def advect_scatter(f, u, v, kind, dt, sem=SEM_GPU):
    f, u, v = _c(f), _c(u), _c(v)
    lib().sfo_advect_scatter(_p(f), _p(u), _p(v), f.shape[0], f.shape[1], kind, dt, sem)
    return f


def smooth(f):
    f = _c(f)
    lib().sfo_smooth(_p(f), f.shape[0], f.shape[1])
    return f


def project(u, v, iters, sem=SEM_GPU, return_p=False):
    u, v = _c(u), _c(v)
    if return_p:
        p, dv = np.empty_like(u), np.empty_like(u)
        lib().sfo_project(_p(u), _p(v), u.shape[0], u.shape[1], iters, sem, _p(p), _p(dv))
        return u, v, p, dv
    lib().sfo_project(_p(u), _p(v), u.shape[0], u.shape[1], iters, sem, None, None)
    return u, v


def steps(d, sd, diffusion_rate, u, v, su, sv, viscosity, dt, kd, kp, smooth=True, sem=SEM_GPU,
          nsteps=1):
    """`nsteps` full solve() steps with constant sources; returns new (d, u, v)."""
    d, u, v, sd, su, sv = (_c(x) for x in (d, u, v, sd, su, sv))
    lib().sfo_steps(_p(d), _p(sd), diffusion_rate, _p(u), _p(v), _p(su), _p(sv), viscosity, dt,
                    d.shape[0], d.shape[1], kd, kp, int(bool(smooth)), sem, nsteps)
    return d, u, v


def canonical_fields(n, threads=None):
    """(d, u, v, sd, su, sv) of SURVEY.md section 8(d); libm in double, rounded to fp32."""
    arrs = [np.empty((n, n), dtype=np.float32) for _ in range(6)]
    L = lib()
    threads = threads or min(os.cpu_count() or 1, 32, max(1, n // 256))
    bounds = np.linspace(0, n, threads + 1).astype(int)

    def work(k):
        L.sfo_canonical_fields(n, int(bounds[k]), int(bounds[k + 1]), *[_p(a) for a in arrs])

    ts = [threading.Thread(target=work, args=(k,)) for k in range(threads)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    return tuple(arrs)


def fnv1a64(a):
    a = np.ascontiguousarray(a)
    return int(lib().sfo_fnv1a64(a.ctypes.data_as(C.c_void_p), a.nbytes))


def field_stats(a):
    a64 = a.astype(np.float64)
    return dict(sum=float(a64.sum()), l2=float(np.sqrt((a64 * a64).sum())), min=float(a.min()),
                max=float(a.max()), fnv="%016x" % fnv1a64(a))
