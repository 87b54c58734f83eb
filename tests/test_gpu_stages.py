"""GPU parity, stage by stage: every stage of the CUDA path (through the C ABI) against the oracle
(SFO_SEM_GPU) on the same seeded inputs.  Bit-exact wherever the reference is deterministic; stated
tolerances for the atomic scatter and for the fp32-corrected diffuse divide."""
import numpy as np
import pytest

from util import DIFFUSION_RATE, DT, VISCOSITY, assert_bitwise, assert_close, rng_fields

pytestmark = pytest.mark.gpu

D, U, V = 0, 1, 2
NAIVE, STREAM = 0, 1
DIV_F64, DIV_F32 = 0, 1

# (jacobi_mode, temporal_block)
RELAX_MODES = [(NAIVE, 1), (STREAM, 1), (STREAM, 2), (STREAM, 4), (STREAM, 8)]
RELAX_IDS = ["naive", "stream1", "stream2", "stream4", "stream8"]


def make(f2d, n, mode=STREAM, T=8, div=DIV_F64, **kw):
    return f2d.FluidSolverB200(n, n, jacobi_mode=mode, temporal_block=T, divide_mode=div, **kw)


@pytest.mark.parametrize("n", [5, 16, 37, 64, 132])
def test_set_bnd_and_add_sources_bitwise(f2d, sfo, gpu_ok, n):
    d, u, v, sd, su, sv = rng_fields(n, 300 + n)
    with make(f2d, n, mode=NAIVE, T=1) as s:
        for kind in (0, 1, 2):
            s.upload(d, u, v)
            s.stage_set_bnd(U, kind)
            assert_bitwise(s.download()[1], sfo.set_bnd(u, kind), "set_bnd kind %d" % kind)
        s.upload(d, u, v)
        s.set_sources(sd, su, sv)
        for f, (a, src) in enumerate(((d, sd), (u, su), (v, sv))):
            s.stage_add_sources(f, DT)
            assert_bitwise(s.download()[f], sfo.add_sources(a, src, DT), "add_sources field %d" % f)


@pytest.mark.parametrize("mode,T", RELAX_MODES, ids=RELAX_IDS)
@pytest.mark.parametrize("n", [8, 64, 132, 260])
def test_diffuse_exact_divide_bitwise(f2d, sfo, gpu_ok, mode, T, n):
    d, u, v, *_ = rng_fields(n, 400 + n)
    with make(f2d, n, mode, T, DIV_F64) as s:
        for kind, rate, iters in ((0, DIFFUSION_RATE, 15), (1, VISCOSITY, 20), (2, 1e-4, 7), (0, 1e-4, 1), (2, DIFFUSION_RATE, 33)):
            s.upload(d, u, v)
            s.stage_diffuse(D, kind, rate, DT, iters)
            assert_bitwise(s.download()[0], sfo.diffuse(d, kind, rate, DT, iters),
                           "diffuse n=%d kind=%d rate=%g K=%d" % (n, kind, rate, iters))


@pytest.mark.parametrize("mode,T", [(NAIVE, 1), (STREAM, 8)], ids=["naive", "stream8"])
def test_diffuse_fp32_corrected_divide_within_1ulp(f2d, sfo, gpu_ok, mode, T):
    """F2D_DIV_F32_CORR replaces the reference's fp64 divide; stated tolerance: <= 2 ulp per cell,
    rel-L2 <= 1e-7, and at most 1 cell in 10^4 may differ at all."""
    n = 256
    d, u, v, *_ = rng_fields(n, 77)
    with make(f2d, n, mode, T, DIV_F32) as s:
        for kind, rate in ((0, DIFFUSION_RATE), (1, VISCOSITY), (2, 1e-4)):
            s.upload(d, u, v)
            s.stage_diffuse(U, kind, rate, DT, 20)
            assert_close(s.download()[1], sfo.diffuse(u, kind, rate, DT, 20), "diffuse rate %g" % rate,
                         max_ulp=2, rel_l2=1e-7, max_frac=1e-4)


@pytest.mark.parametrize("mode,T", [(NAIVE, 1), (STREAM, 1), (STREAM, 8)], ids=["naive", "stream1", "stream8"])
@pytest.mark.parametrize("a", [1.7e5, 2.7e6, 1.07e7, 4.3e7])
def test_diffuse_fp32_corrected_divide_at_the_published_coefficients(f2d, sfo, gpu_ok, mode, T, a):
    """The diffusion coefficient a = dt * N^2 * rate of the configurations the numbers are published for:
    1.7e5 (4096^2), 2.7e6 (16384^2), 1.07e7 (32768^2; c = 1 + 4a no longer fits 24 bits; the rc + rl split of the
    reciprocal carries 48) and 4.3e7 beyond.  f2d_stage_diffuse takes the rate, so a 256^2 grid reproduces each of
    them.  Against the reference's fp64 divide after 20 sweeps: <= 2 ulp per cell, rel-L2 <= 1e-7, at most 1 cell
    in 10^4 different at all."""
    n = 256
    rate = a / (DT * n * n)
    assert abs(float(sfo.lib().sfo_diffuse_coeff(n, n, rate, DT)) / a - 1) < 1e-3
    d, u, v, *_ = rng_fields(n, 78)
    with make(f2d, n, mode, T, DIV_F32) as s:
        for fld, kind in ((D, 0), (U, 1), (V, 2)):
            s.upload(d, u, v)
            s.stage_diffuse(fld, kind, rate, DT, 20)
            e = assert_close(s.download()[fld], sfo.diffuse((d, u, v)[fld], kind, rate, DT, 20), "a=%g kind=%d" % (a, kind),
                             max_ulp=2, rel_l2=1e-7, max_frac=1e-4)
            print("fp32-corrected divide a=%g kind=%d T=%d: %r" % (a, kind, T, e))


@pytest.mark.parametrize("mode,T", RELAX_MODES, ids=RELAX_IDS)
@pytest.mark.parametrize("n", [8, 64, 132, 260])
def test_project_bitwise(f2d, sfo, gpu_ok, mode, T, n):
    d, u, v, *_ = rng_fields(n, 500 + n)
    with make(f2d, n, mode, T) as s:
        for iters in (1, 5, 20, 43):
            s.upload(d, u, v)
            s.stage_project(iters)
            gd, gu, gv = s.download()
            ou, ov, op, odv = sfo.project(u, v, iters, return_p=True)
            assert_bitwise(s.download_field(7), odv, "divergence n=%d K=%d" % (n, iters))
            assert_bitwise(s.download_field(6), op, "pressure n=%d K=%d" % (n, iters))
            assert_bitwise(gu, ou, "project u n=%d K=%d" % (n, iters))
            assert_bitwise(gv, ov, "project v n=%d K=%d" % (n, iters))


@pytest.mark.parametrize("n,cells", [(16, 2.0), (64, 4.0), (132, 9.0), (37, 40.0)])
def test_advect_velocity_gather_bitwise(f2d, sfo, gpu_ok, n, cells):
    """cells=40 on a 37-cell grid drives most back-traces into the [1.5, N-1.5] clamp."""
    d, u, v, *_ = rng_fields(n, 600 + n, vel_cells=cells)
    with make(f2d, n, mode=NAIVE, T=1) as s:
        s.upload(d, u, v)
        s.stage_advect_velocity(DT)
        _, gu, gv = s.download()
        assert_bitwise(gu, sfo.advect_gather(u, u, v, 1, DT), "advect u")
        assert_bitwise(gv, sfo.advect_gather(v, u, v, 2, DT), "advect v")


@pytest.mark.parametrize("n,cells", [(16, 2.0), (64, 4.0), (132, 9.0), (37, 40.0)])
def test_advect_density_scatter_within_summation_order(f2d, sfo, gpu_ok, n, cells):
    """The reference scatters with float atomics in hardware order (gpu.cu:156-159), so bit-equality
    is undefined; stated tolerance: max-abs <= 1e-6 * max(1,|f|max) (a few ulp), rel-L2 <= 5e-7."""
    d, u, v, *_ = rng_fields(n, 700 + n, vel_cells=cells)
    with make(f2d, n, mode=NAIVE, T=1) as s:
        s.upload(d, u, v)
        s.stage_advect_density(DT)
        s.sync()
        assert_close(s.download()[0], sfo.advect_scatter(d, u, v, 0, DT), "scatter n=%d" % n,
                     max_abs_rel=1e-6, rel_l2=5e-7)


@pytest.mark.parametrize("n", [5, 64, 37])
def test_smooth_bitwise(f2d, sfo, gpu_ok, n):
    d, u, v, *_ = rng_fields(n, 800 + n)
    with make(f2d, n, mode=NAIVE, T=1) as s:
        s.upload(d, u, v)
        s.stage_smooth()
        assert_bitwise(s.download()[0], sfo.smooth(d), "smooth")


def test_temporal_blocking_equals_single_sweeps_at_full_size(f2d, gpu_ok):
    """Size-independent property at the roofline size (4096^2, 80 sweeps): T fused sweeps are
    bit-identical to T single sweeps of the naive kernel, for diffuse and for the pressure solve."""
    n = 4096
    d, u, v, *_ = rng_fields(n, 9)
    res = {}
    for tag, (mode, T) in (("naive", (NAIVE, 1)), ("stream8", (STREAM, 8)), ("stream4", (STREAM, 4))):
        with make(f2d, n, mode, T) as s:
            s.upload(d, u, v)
            s.stage_diffuse(D, 0, DIFFUSION_RATE, DT, 80)
            s.stage_diffuse(U, 1, VISCOSITY, DT, 80)
            s.stage_project(80)
            res[tag] = s.download() + (s.download_field(6),)
    for tag in ("stream8", "stream4"):
        for name, a, b in zip(("d", "u", "v", "p"), res[tag], res["naive"]):
            assert_bitwise(a, b, "%s vs naive: %s" % (tag, name))
