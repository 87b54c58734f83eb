"""GPU parity of the fluid_solver_cpu-compatible mode (F2D_SEM_CPU).

Every stage and the full step, through the C ABI, against the oracle's SFO_SEM_CPU restatement -- which
tests/test_oracle_pin.py pins bit for bit to the UNMODIFIED fluid_solver_cpu and to the FNV anchors of SURVEY.md
Appendix D.  Everything here is BIT-EXACT, the density included: the Gauss-Seidel sweeps run as a dependency-
respecting wavefront (src/fluid_solver_cpu.cpp:104-113, :196-204) and the scatter adds in the reference's
lexicographic source order (:127-152)."""
import json
import os

import numpy as np
import pytest

from util import DIFFUSION_RATE, DT, VISCOSITY, assert_bitwise, rng_fields

pytestmark = pytest.mark.gpu
D, U, V = 0, 1, 2
SEM_CPU = 1
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def make(f2d, rows, cols=None, **kw):
    return f2d.FluidSolverB200(rows, cols or rows, semantics=SEM_CPU, smooth=False, **kw)


@pytest.mark.parametrize("n", [5, 16, 37, 64, 132])
def test_set_bnd_and_add_sources(f2d, sfo, gpu_ok, n):
    d, u, v, sd, su, sv = rng_fields(n, 1300 + n)
    with make(f2d, n) as s:
        for kind in (0, 1, 2):
            s.upload(d, u, v)
            s.stage_set_bnd(U, kind)
            assert_bitwise(s.download()[1], sfo.set_bnd(u, kind, sfo.SEM_CPU), "set_bnd kind %d" % kind)
        s.upload(d, u, v)
        s.set_sources(sd, su, sv)
        for f, (a, src) in enumerate(((d, sd), (u, su), (v, sv))):
            s.stage_add_sources(f, DT)
            assert_bitwise(s.download()[f], sfo.add_sources(a, src, DT, sfo.SEM_CPU), "add_sources field %d" % f)


@pytest.mark.parametrize("n", [3, 8, 33, 34, 35, 64, 67, 100, 132, 260])
def test_gauss_seidel_diffuse(f2d, sfo, gpu_ok, n):
    d, u, v, *_ = rng_fields(n, 1400 + n)
    with make(f2d, n) as s:
        for kind, rate, iters in ((0, DIFFUSION_RATE, 20), (1, VISCOSITY, 20), (2, 1e-4, 7), (0, 1e-4, 1), (2, DIFFUSION_RATE, 33)):
            s.upload(d, u, v)
            s.stage_diffuse(U, kind, rate, DT, iters)
            s.sync()
            assert_bitwise(s.download()[1], sfo.diffuse(u, kind, rate, DT, iters, sfo.SEM_CPU),
                           "GS diffuse n=%d kind=%d rate=%g K=%d" % (n, kind, rate, iters))


def test_gauss_seidel_diffuse_many_sweeps_more_warps_than_resident(f2d, sfo, gpu_ok):
    """520^2 x 200 sweeps = 3400 warps per problem: more than fit on the GPU at once, so late sweeps start only
    when early ones retire (the ticket order keeps the wavefront deadlock-free)."""
    n, iters = 520, 200
    d, u, v, *_ = rng_fields(n, 1500)
    with make(f2d, n) as s:
        s.upload(d, u, v)
        s.stage_diffuse(D, 0, DIFFUSION_RATE, DT, iters)
        s.sync()
        assert_bitwise(s.download()[0], sfo.diffuse(d, 0, DIFFUSION_RATE, DT, iters, sfo.SEM_CPU), "GS diffuse 520^2 K=200")


@pytest.mark.parametrize("rows,cols", [(40, 100), (130, 36)])
def test_gauss_seidel_diffuse_non_square(f2d, sfo, gpu_ok, rows, cols):
    """The reference's grid<T>::cols() bug (src/grid.hpp:20-22) limits IT to square grids; the oracle and the
    kernels are general."""
    d, u, v, *_ = rng_fields(rows, 1600 + rows, cols=cols)
    with make(f2d, rows, cols) as s:
        s.upload(d, u, v)
        s.stage_diffuse(V, 2, DIFFUSION_RATE, DT, 11)
        s.sync()
        assert_bitwise(s.download()[2], sfo.diffuse(v, 2, DIFFUSION_RATE, DT, 11, sfo.SEM_CPU), "GS diffuse %dx%d" % (rows, cols))


@pytest.mark.parametrize("n,iters", [(8, 3), (64, 20), (100, 20), (132, 40), (260, 20), (64, 0)])
def test_project(f2d, sfo, gpu_ok, n, iters):
    d, u, v, *_ = rng_fields(n, 1700 + n)
    with make(f2d, n) as s:
        s.upload(d, u, v)
        s.stage_project(iters)
        s.sync()
        _, gu, gv = s.download()
        gp, gdv = s.download_field(f2d.FIELD_PRESSURE), s.download_field(f2d.FIELD_DIVERGENCE)
    ou, ov, op, odv = sfo.project(u, v, iters, sfo.SEM_CPU, return_p=True)
    assert_bitwise(gdv, odv, "divergence")
    assert_bitwise(gp, op, "pressure")
    assert_bitwise(gu, ou, "u")
    assert_bitwise(gv, ov, "v")


@pytest.mark.parametrize("n,cells", [(16, 2.0), (64, 3.0), (100, 6.5), (132, 0.3), (96, 40.0)])
def test_advect_velocity_and_ordered_density_scatter(f2d, sfo, gpu_ok, n, cells):
    d, u, v, *_ = rng_fields(n, 1800 + n, vel_cells=cells)
    with make(f2d, n) as s:
        s.upload(d, u, v)
        s.stage_advect_velocity(DT)
        _, gu, gv = s.download()
        s.upload(d, u, v)
        s.stage_advect_density(DT)
        gd = s.download()[0]
    assert_bitwise(gu, sfo.advect_gather(u, u, v, 1, DT, sfo.SEM_CPU), "advect u")
    assert_bitwise(gv, sfo.advect_gather(v, u, v, 2, DT, sfo.SEM_CPU), "advect v")
    # deterministic: the additions happen in the CPU solver's source order
    assert_bitwise(gd, sfo.advect_scatter(d, u, v, 0, DT, sfo.SEM_CPU), "ordered scatter (max displacement %.1f cells)" % cells)


@pytest.mark.parametrize("n,kd,kp,steps", [(64, 20, 20, 3), (100, 7, 9, 3), (128, 20, 20, 2), (37, 5, 0, 2)])
@pytest.mark.parametrize("graph", [True, False], ids=["graph", "eager"])
def test_full_step_bit_identical_to_cpu_solver_semantics(f2d, sfo, gpu_ok, n, kd, kp, steps, graph):
    f = rng_fields(n, 1900 + n)
    with make(f2d, n, diffuse_iters=kd, project_iters=kp, use_graph=graph) as s:
        s.upload(*f[:3])
        s.set_sources(*f[3:])
        s.step(DIFFUSION_RATE, VISCOSITY, DT, steps)
        s.sync()
        gd, gu, gv = s.download()
    od, ou, ov = sfo.steps(f[0], f[3], DIFFUSION_RATE, f[1], f[2], f[4], f[5], VISCOSITY, DT, kd, kp, smooth=False,
                           sem=sfo.SEM_CPU, nsteps=steps)
    assert_bitwise(gu, ou, "u")
    assert_bitwise(gv, ov, "v")
    assert_bitwise(gd, od, "d")


@pytest.mark.parametrize("steps", [1, 10, 100])
def test_reproduces_the_anchors_of_the_unmodified_cpu_solver(f2d, sfo, gpu_ok, steps):
    """FNV-1a-64 hashes of d, u, v after 1 / 10 / 100 steps of the UNMODIFIED
    fluid_solver_cpu::solve on the canonical 256^2 input -- reproduced here by the GPU, hash for hash."""
    anchors = json.load(open(os.path.join(GOLDEN, "anchors_ref_cpu_256.json")))
    d, u, v, sd, su, sv = sfo.canonical_fields(256)
    with f2d.FluidSolverB200.cpu_compatible(256, 256) as s:
        s.upload(d, u, v)
        s.set_sources(sd, su, sv)
        s.step(DIFFUSION_RATE, VISCOSITY, DT, steps)
        s.sync()
        fields = s.download()
    for name, a in zip("duv", fields):
        assert "%016x" % sfo.fnv1a64(a) == anchors[str(steps)][name]["fnv"], (name, steps)


def test_full_size_step_2048(f2d, sfo, gpu_ok):
    """One fluid_solver_cpu::solve step at 2048^2 on the canonical fields (the oracle needs a few seconds for it):
    64 row bands x 64 column tiles x 20 sweeps x 3 problems in flight, every field bit-identical."""
    n = 2048
    d, u, v, sd, su, sv = sfo.canonical_fields(n)
    with f2d.FluidSolverB200.cpu_compatible(n, n) as s:
        s.upload(d, u, v)
        s.set_sources(sd, su, sv)
        s.step(DIFFUSION_RATE, VISCOSITY, DT, 1)
        s.sync()
        gd, gu, gv = s.download()
    od, ou, ov = sfo.steps(d, sd, DIFFUSION_RATE, u, v, su, sv, VISCOSITY, DT, 20, 20, smooth=False, sem=sfo.SEM_CPU, nsteps=1)
    assert_bitwise(gu, ou, "u")
    assert_bitwise(gv, ov, "v")
    assert_bitwise(gd, od, "d")


def test_solve_host_against_live_reference_cpu_solver(f2d, sfo, gpu_ok):
    """fluid_solver::solve through the C ABI on host grids vs the unmodified fluid_solver_cpu::solve
    (oracle/_ref/libref_cpu.so travels to the box; falls back to the pinned oracle when it did not)."""
    from oracle import refs

    n = 160
    d, u, v, sd, su, sv = rng_fields(n, 2024)
    hd, hu, hv = d.copy(), u.copy(), v.copy()
    with f2d.FluidSolverB200.cpu_compatible(n, n) as s:
        for _ in range(2):
            s.solve(hd, sd, DIFFUSION_RATE, hu, hv, su, sv, VISCOSITY, DT)
    if refs.have_cpu():
        rd, ru, rv = refs.ref_cpu().solve(d, sd, DIFFUSION_RATE, u, v, su, sv, VISCOSITY, DT, 2)
    else:
        rd, ru, rv = sfo.steps(d, sd, DIFFUSION_RATE, u, v, su, sv, VISCOSITY, DT, 20, 20, smooth=False, sem=sfo.SEM_CPU, nsteps=2)
    assert_bitwise(hu, ru, "u")
    assert_bitwise(hv, rv, "v")
    assert_bitwise(hd, rd, "d")


def test_cpu_semantics_is_single_gpu_only(f2d, gpu_ok):
    with pytest.raises(f2d.F2DError):
        f2d.FluidSolverB200(64, 64, semantics=SEM_CPU, global_rows=128, row_offset=0, halo=8)
