"""ctypes bindings of the UNMODIFIED reference solvers (oracle/_ref).  TEST INFRASTRUCTURE ONLY.

libref_cpu.so / libref_gpu.so are produced by oracle/Makefile from the sources under
/root/reference (never copied into this repo) plus the extern "C" shims oracle/ref_*_shim.*.
They are git-ignored but travel to the GPU box with the gpurun snapshot.  Square grids only
(the reference's grid<T>::cols() returns rows, src/grid.hpp:20-22).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
CPU_PATH = os.path.join(_HERE, "_ref", "libref_cpu.so")
GPU_PATH = os.path.join(_HERE, "_ref", "libref_gpu.so")
RENDER_PATH = os.path.join(_HERE, "_ref", "libref_render.so")
_FP = C.POINTER(C.c_float)
_cpu = None
_gpu = None


def have_cpu():
    return os.path.exists(CPU_PATH)


def have_gpu():
    """True when libref_gpu.so exists AND a CUDA device is visible."""
    if not os.path.exists(GPU_PATH):
        return False
    try:
        return gpu().ref_gpu_device_count() > 0
    except OSError:
        return False


def cpu():
    global _cpu
    if _cpu is None:
        L = C.CDLL(CPU_PATH)
        sz, f, i, u = C.c_size_t, C.c_float, C.c_int, C.c_uint
        L.ref_cpu_solve.argtypes = [sz, _FP, _FP, f, _FP, _FP, _FP, _FP, f, f, u]
        L.ref_cpu_step_k.argtypes = [sz, _FP, _FP, f, _FP, _FP, _FP, _FP, f, f, u, u, u]
        L.ref_cpu_set_bnd.argtypes = [sz, _FP, i]
        L.ref_cpu_add_sources.argtypes = [sz, _FP, _FP, f]
        L.ref_cpu_diffuse.argtypes = [sz, _FP, i, f, f, u]
        L.ref_cpu_advect.argtypes = [sz, _FP, _FP, _FP, i, f, i]
        L.ref_cpu_project.argtypes = [sz, _FP, _FP, u]
        for n in ("ref_cpu_solve", "ref_cpu_step_k", "ref_cpu_set_bnd", "ref_cpu_add_sources",
                  "ref_cpu_diffuse", "ref_cpu_advect", "ref_cpu_project"):
            getattr(L, n).restype = None
        _cpu = L
    return _cpu


def gpu():
    global _gpu
    if _gpu is None:
        L = C.CDLL(GPU_PATH)
        sz, f, i, u = C.c_size_t, C.c_float, C.c_int, C.c_uint
        L.ref_gpu_device_count.restype = C.c_int
        L.ref_gpu_solve.argtypes = [sz, _FP, _FP, f, _FP, _FP, _FP, _FP, f, f, u]
        L.ref_gpu_solve.restype = C.c_double
        L.ref_gpu_step_k.argtypes = [sz, _FP, _FP, f, _FP, _FP, _FP, _FP, f, f, u, u, i, u]
        L.ref_gpu_step_k.restype = C.c_double
        L.ref_gpu_set_bnd.argtypes = [sz, _FP, i]
        L.ref_gpu_add_sources.argtypes = [sz, _FP, _FP, f]
        L.ref_gpu_diffuse.argtypes = [sz, _FP, i, f, f, u]
        L.ref_gpu_smooth.argtypes = [sz, _FP]
        L.ref_gpu_advect.argtypes = [sz, _FP, _FP, _FP, i, f, i]
        L.ref_gpu_project.argtypes = [sz, _FP, _FP, u, _FP, _FP]
        for n in ("ref_gpu_set_bnd", "ref_gpu_add_sources", "ref_gpu_diffuse", "ref_gpu_smooth",
                  "ref_gpu_advect", "ref_gpu_project"):
            getattr(L, n).restype = None
        _gpu = L
    return _gpu


def _p(a):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"] and a.shape[0] == a.shape[1]
    return a.ctypes.data_as(_FP)


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float32).copy()


class _Ref:
    """Common stage API over either reference library."""

    def __init__(self, which):
        self.which = which
        self.L = cpu() if which == "cpu" else gpu()
        self.px = "ref_%s_" % which

    def _f(self, name):
        return getattr(self.L, self.px + name)

    def set_bnd(self, f, kind):
        f = _c(f)
        self._f("set_bnd")(f.shape[0], _p(f), kind)
        return f

    def add_sources(self, f, s, dt):
        f, s = _c(f), _c(s)
        self._f("add_sources")(f.shape[0], _p(f), _p(s), dt)
        return f

    def diffuse(self, f, kind, rate, dt, iters):
        f = _c(f)
        self._f("diffuse")(f.shape[0], _p(f), kind, rate, dt, iters)
        return f

    def advect(self, f, u, v, kind, dt, trace):
        f, u, v = _c(f), _c(u), _c(v)
        self._f("advect")(f.shape[0], _p(f), _p(u), _p(v), kind, dt, int(trace))
        return f

    def smooth(self, f):
        assert self.which == "gpu", "only fluid_solver_gpu has smooth (src/fluid_solver_gpu.cu:314)"
        f = _c(f)
        self._f("smooth")(f.shape[0], _p(f))
        return f

    def project(self, u, v, iters, return_p=False):
        u, v = _c(u), _c(v)
        if self.which == "cpu":
            self._f("project")(u.shape[0], _p(u), _p(v), iters)
            return u, v
        if return_p:
            p, dv = np.empty_like(u), np.empty_like(u)
            self._f("project")(u.shape[0], _p(u), _p(v), iters, _p(p), _p(dv))
            return u, v, p, dv
        self._f("project")(u.shape[0], _p(u), _p(v), iters, None, None)
        return u, v

    def solve(self, d, sd, diffusion_rate, u, v, su, sv, viscosity, dt, nsteps=1):
        """The reference's own solve(), literally.  Returns (d,u,v[,ms])."""
        d, u, v, sd, su, sv = (_c(x) for x in (d, u, v, sd, su, sv))
        r = self._f("solve")(d.shape[0], _p(d), _p(sd), diffusion_rate, _p(u), _p(v), _p(su),
                             _p(sv), viscosity, dt, nsteps)
        return (d, u, v) if self.which == "cpu" else (d, u, v, float(r))

    def step_k(self, d, sd, diffusion_rate, u, v, su, sv, viscosity, dt, kd, kp, smooth=True,
               nsteps=1):
        """solve()'s stage sequence with free iteration counts."""
        d, u, v, sd, su, sv = (_c(x) for x in (d, u, v, sd, su, sv))
        if self.which == "cpu":
            self.L.ref_cpu_step_k(d.shape[0], _p(d), _p(sd), diffusion_rate, _p(u), _p(v), _p(su),
                                  _p(sv), viscosity, dt, kd, kp, nsteps)
            return d, u, v
        ms = self.L.ref_gpu_step_k(d.shape[0], _p(d), _p(sd), diffusion_rate, _p(u), _p(v), _p(su),
                                   _p(sv), viscosity, dt, kd, kp, int(bool(smooth)), nsteps)
        return d, u, v, float(ms)


def ref_cpu():
    return _Ref("cpu")


def ref_gpu():
    return _Ref("gpu")


# ---- the reference's renderers (src/density_grid_renderer.cu, src/velocity_grid_renderer.cu), unmodified, compiled
# against oracle/sfml_stub: what the app would hand to SFML for a given field
_render = None


def have_render():
    """True when libref_render.so exists AND a CUDA device is visible."""
    if not os.path.exists(RENDER_PATH):
        return False
    try:
        return render_lib().ref_render_device_count() > 0
    except OSError:
        return False


def render_lib():
    global _render
    if _render is None:
        L = C.CDLL(RENDER_PATH)
        L.ref_render_device_count.restype = C.c_int
        L.ref_render_density.argtypes = [C.c_size_t, _FP, C.c_float, C.c_float, C.c_float, C.c_uint, C.c_uint,
                                         C.POINTER(C.c_ubyte)]
        L.ref_render_density.restype = C.c_int
        L.ref_render_velocity.argtypes = [C.c_size_t, _FP, _FP, C.c_uint, C.c_uint, _FP, C.POINTER(C.c_int)]
        L.ref_render_velocity.restype = C.c_int
        _render = L
    return _render


def ref_render_density(d, mult, target=(800, 800)):
    """density_grid_renderer::draw: (n, n, 4) uint8 RGBA."""
    d = _c(d)
    n = d.shape[0]
    img = np.empty((n, n, 4), np.uint8)
    rc = render_lib().ref_render_density(n, _p(d), mult[0], mult[1], mult[2], target[0], target[1],
                                         img.ctypes.data_as(C.POINTER(C.c_ubyte)))
    assert rc == 0, rc
    return img


def ref_render_velocity(u, v, target=(800, 800)):
    """velocity_grid_renderer::draw: (n, n, 4) float32 segments (start.x, start.y, end.x, end.y); the scales are
    target / grid as the reference computes them (src/velocity_grid_renderer.cu:64)."""
    u, v = _c(u), _c(v)
    n = u.shape[0]
    ln = np.empty((n, n, 4), np.float32)
    ok = C.c_int(0)
    rc = render_lib().ref_render_velocity(n, _p(u), _p(v), target[0], target[1], ln.ctypes.data_as(_FP), C.byref(ok))
    assert rc == 0 and ok.value == 1, (rc, ok.value)
    return ln
