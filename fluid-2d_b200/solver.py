"""Host-side mirror of the reference's solver interface over the C ABI.

`FluidSolverB200.solve` has the argument list and in-place semantics of
`fluid_solver::solve` (src/fluid_solver.hpp:16-24): six host grids + three scalars, density / u / v
updated in place, sources untouched.  The constructor mirrors `fluid_solver_gpu(rows, cols)`
(src/fluid_solver_gpu.cu:209-218).  Everything else (device-resident stepping, single stages) is
the extension surface of include/f2d.h used by the parity tests and bench.py.
"""
import ctypes as C

import numpy as np

from . import capi

_FP = C.POINTER(C.c_float)


def _host(a, writable=False):
    if not isinstance(a, np.ndarray) or a.dtype != np.float32 or not a.flags["C_CONTIGUOUS"]:
        raise TypeError("grids must be C-contiguous float32 numpy arrays (like grid<float>::data())")
    if writable and not a.flags["WRITEABLE"]:
        raise TypeError("grid is updated in place and must be writable")
    return a.ctypes.data_as(_FP)


class FluidSolverB200:
    """Drop-in for fluid_solver_gpu on one B200 (or one row slab of a multi-GPU run).

    `semantics=capi.SEM_CPU` switches the arithmetic to fluid_solver_cpu's (in-place Gauss-Seidel, no FMA,
    averaged corners, ordered scatter; src/fluid_solver_cpu.cpp): with diffuse_iters=project_iters=20 and
    smooth=False the results are bit-identical to fluid_solver_cpu::solve.  See `cpu_compatible`."""

    @classmethod
    def cpu_compatible(cls, rows, cols, iters=20, **kw):
        """The solver configured as fluid_solver_cpu::solve (src/fluid_solver_cpu.cpp:15-30)."""
        return cls(rows, cols, diffuse_iters=iters, project_iters=iters, smooth=False, semantics=capi.SEM_CPU, **kw)

    def __init__(self, rows, cols, diffuse_iters=15, project_iters=20, smooth=True, jacobi_mode=None,
                 temporal_block=0, temporal_block_diffuse=0, divide_mode=capi.DIV_F32_CORR, use_graph=True, device=-1,
                 global_rows=None, row_offset=0, halo=0, stream=None, semantics=capi.SEM_GPU):
        self._h = C.c_void_p()
        L = capi.load()
        cfg = capi.SolverConfig()
        capi.check(L.f2d_config_default(C.byref(cfg), rows, cols))
        cfg.diffuse_iters = diffuse_iters
        cfg.project_iters = project_iters
        cfg.smooth = 1 if smooth else 0
        if jacobi_mode is not None:
            cfg.jacobi_mode = jacobi_mode
        cfg.temporal_block = temporal_block
        cfg.temporal_block_diffuse = temporal_block_diffuse
        cfg.divide_mode = divide_mode
        cfg.use_graph = 1 if use_graph else 0
        cfg.device = device
        cfg.global_rows = rows if global_rows is None else global_rows
        cfg.row_offset = row_offset
        cfg.halo = halo
        cfg.stream = stream
        cfg.semantics = semantics
        capi.check(L.f2d_create(C.byref(cfg), C.byref(self._h)))
        self._L = L
        self.rows, self.cols = rows, cols

    # ---- lifetime
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._L.f2d_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _shape_ok(self, *arrs):
        for a in arrs:
            if a is not None and a.shape != (self.rows, self.cols):
                raise ValueError("grid shape %r does not match solver (%d, %d)" % (a.shape, self.rows, self.cols))

    # ---- the reference interface (src/fluid_solver.hpp:16-24)
    def solve(self, density, density_source, diffusion_rate, u, v, u_source, v_source, viscosity, dt):
        self._shape_ok(density, density_source, u, v, u_source, v_source)
        capi.check(self._L.f2d_solve_host(self._h, _host(density, True), _host(density_source), diffusion_rate,
                                          _host(u, True), _host(v, True), _host(u_source), _host(v_source),
                                          viscosity, dt))

    def pin(self, *arrays):
        """Page-lock caller-owned grids (f2d_pin_host) so that solve() moves them by DMA.  Each array must stay alive
        and in place until unpin() or close(); solve() never pins behind the caller's back."""
        for a in arrays:
            capi.check(self._L.f2d_pin_host(self._h, a.ctypes.data_as(C.c_void_p), a.nbytes))

    def unpin(self, *arrays):
        for a in arrays:
            capi.check(self._L.f2d_unpin_host(self._h, a.ctypes.data_as(C.c_void_p)))

    # ---- device-resident extension
    def upload(self, density=None, u=None, v=None):
        self._shape_ok(density, u, v)
        capi.check(self._L.f2d_upload(self._h, *[None if a is None else _host(a) for a in (density, u, v)]))

    def set_sources(self, density_source=None, u_source=None, v_source=None):
        self._shape_ok(density_source, u_source, v_source)
        capi.check(self._L.f2d_set_sources(
            self._h, *[None if a is None else _host(a) for a in (density_source, u_source, v_source)]))

    def clear_sources(self):
        capi.check(self._L.f2d_clear_sources(self._h))

    def download(self):
        d, u, v = (np.empty((self.rows, self.cols), np.float32) for _ in range(3))
        capi.check(self._L.f2d_download(self._h, _host(d, True), _host(u, True), _host(v, True)))
        return d, u, v

    def upload_field(self, field, a):
        self._shape_ok(a)
        capi.check(self._L.f2d_upload_field(self._h, field, _host(a)))

    def download_field(self, field):
        a = np.empty((self.rows, self.cols), np.float32)
        capi.check(self._L.f2d_download_field(self._h, field, _host(a, True)))
        return a

    def step(self, diffusion_rate, viscosity, dt, nsteps=1):
        capi.check(self._L.f2d_step(self._h, diffusion_rate, viscosity, dt, nsteps))

    def step_timed(self, diffusion_rate, viscosity, dt, nsteps=1):
        ms = C.c_float()
        capi.check(self._L.f2d_step_timed(self._h, diffusion_rate, viscosity, dt, nsteps, C.byref(ms)))
        return float(ms.value)

    def sync(self):
        capi.check(self._L.f2d_sync(self._h))

    # ---- single stages (parity tests)
    def stage_set_bnd(self, field, kind):
        capi.check(self._L.f2d_stage_set_bnd(self._h, field, kind))

    def stage_add_sources(self, field, dt):
        capi.check(self._L.f2d_stage_add_sources(self._h, field, dt))

    def stage_diffuse(self, field, kind, rate, dt, iters):
        capi.check(self._L.f2d_stage_diffuse(self._h, field, kind, rate, dt, iters))

    def stage_smooth(self):
        capi.check(self._L.f2d_stage_smooth(self._h))

    def stage_advect_density(self, dt):
        capi.check(self._L.f2d_stage_advect_density(self._h, dt))

    def stage_advect_velocity(self, dt):
        capi.check(self._L.f2d_stage_advect_velocity(self._h, dt))

    def stage_project(self, iters):
        capi.check(self._L.f2d_stage_project(self._h, iters))

    # ---- headless renderers
    def render_density_rgba(self, mult=(255.0, 255.0, 255.0)):
        """density -> RGBA8 image (rows, cols, 4), as grid_to_image_kernel (src/density_grid_renderer.cu:10-29)."""
        img = np.empty((self.rows, self.cols, 4), np.uint8)
        capi.check(self._L.f2d_render_density_rgba(self._h, mult[0], mult[1], mult[2], img.ctypes.data_as(C.POINTER(C.c_ubyte))))
        return img

    def render_velocity_lines(self, horizontal_scale=1.0, vertical_scale=1.0):
        """(rows, cols, 4) segments (start.x, start.y, end.x, end.y), as velocity_to_lines_kernel
        (src/velocity_grid_renderer.cu:8-44)."""
        ln = np.empty((self.rows, self.cols, 4), np.float32)
        capi.check(self._L.f2d_render_velocity_lines(self._h, horizontal_scale, vertical_scale, ln.ctypes.data_as(_FP)))
        return ln

    # ---- measurement / interop
    def bench_jacobi(self, diffuse_like, iters, reps):
        ms = C.c_float()
        capi.check(self._L.f2d_bench_jacobi(self._h, int(diffuse_like), iters, reps, C.byref(ms)))
        return float(ms.value)

    def launch_count(self):
        n = C.c_uint64()
        capi.check(self._L.f2d_launch_count(self._h, C.byref(n)))
        return int(n.value)

    def field_ptr(self, field):
        p, pitch = C.c_void_p(), C.c_size_t()
        capi.check(self._L.f2d_field_ptr(self._h, field, C.byref(p), C.byref(pitch)))
        return int(p.value), int(pitch.value)

    def config(self):
        cfg = capi.SolverConfig()
        capi.check(self._L.f2d_get_config(self._h, C.byref(cfg)))
        return cfg

    def stream(self):
        p = C.c_void_p()
        capi.check(self._L.f2d_get_stream(self._h, C.byref(p)))
        return int(p.value or 0)
