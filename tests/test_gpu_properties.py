"""Size-independent properties of the CUDA path at the published single-GPU size (4096^2, K = 80), where the oracle
would need minutes per stage: exact linearity under power-of-two scaling, identities, and fused == unfused with the
product's default arithmetic.  All through the C ABI, all bit-exact (scaling by a power of two commutes with every
rounding of the step as long as nothing under- or overflows, which these O(1) fields do not)."""
import numpy as np
import pytest

from util import DIFFUSION_RATE, DT, VISCOSITY, assert_bitwise, rng_fields

pytestmark = pytest.mark.gpu

D, U, V, P = 0, 1, 2, 6
NAIVE, STREAM = 0, 1
DIV_F64, DIV_F32 = 0, 1
N, K = 4096, 80


def make(f2d, mode=STREAM, T=8, div=DIV_F32, **kw):
    return f2d.FluidSolverB200(N, N, jacobi_mode=mode, temporal_block=T, divide_mode=div, **kw)


@pytest.fixture(scope="module")
def fields():
    return rng_fields(N, 4242)


def relax_all(s, d, u, v):
    """80 diffuse sweeps on d (rate 0.5) and u (viscosity), then project(80): returns d, u, v, p"""
    s.upload(d, u, v)
    s.stage_diffuse(D, 0, DIFFUSION_RATE, DT, K)
    s.stage_diffuse(U, 1, VISCOSITY, DT, K)
    s.stage_project(K)
    return s.download() + (s.download_field(P),)


@pytest.mark.parametrize("scale", [4.0, 0.125])
def test_relaxations_are_exactly_linear_under_power_of_two_scaling(f2d, gpu_ok, fields, scale):
    """diffuse (default two-operation constant division), divergence, the pressure solve with its 4^s-scaled levels and
    the gradient subtract are linear maps whose every rounding commutes with a power-of-two factor: relax(c * f) must
    equal c * relax(f) bit for bit."""
    d, u, v = fields[:3]
    c = np.float32(scale)
    with make(f2d) as s:
        base = relax_all(s, d, u, v)
        scaled = relax_all(s, d * c, u * c, v * c)
    for name, a, b in zip("duvp", scaled, base):
        assert_bitwise(a, b * c, "relax(%g * f) vs %g * relax(f): %s" % (scale, scale, name))


def test_default_arithmetic_fused_equals_single_sweeps(f2d, gpu_ok, fields):
    """The product's defaults (fp32 constant division, scaled pressure levels, level lag, T = 8) against one naive
    sweep per launch with the same division: bit-identical after 80 sweeps."""
    d, u, v = fields[:3]
    with make(f2d, STREAM, 8) as s:
        fused = relax_all(s, d, u, v)
    with make(f2d, NAIVE, 1) as s:
        single = relax_all(s, d, u, v)
    for name, a, b in zip("duvp", fused, single):
        assert_bitwise(a, b, "T=8 vs single sweeps: %s" % name)


def test_diffuse_with_rate_zero_is_the_identity(f2d, gpu_ok, fields):
    """a = 0: every sweep returns x0 on the interior (c = 1, rc = 1, rl = 0) and the boundary pass mirrors it."""
    d, u, v = fields[:3]
    with make(f2d) as s:
        s.upload(d, u, v)
        s.stage_diffuse(D, 0, 0.0, DT, K)
        out = s.download()[0]
    assert_bitwise(out[1:-1, 1:-1], d[1:-1, 1:-1], "interior")
    assert_bitwise(out[0, 1:-1], d[1, 1:-1], "top edge = first interior row")
    assert_bitwise(out[1:-1, -1], d[1:-1, -2], "right edge = last interior column")
    for i, j in ((0, 0), (0, -1), (-1, 0), (-1, -1)):
        assert out[i, j] == d[i, j], "corners are never written"


def test_density_advection_by_zero_velocity_is_the_identity(f2d, gpu_ok, fields):
    """u = v = 0: every source lands on its own cell with weight exactly 1 (one atomic add per target, so the result
    does not depend on the order of the float atomics either)."""
    d = fields[0]
    z = np.zeros_like(d)
    with make(f2d, smooth=False) as s:
        s.upload(d, z, z)
        s.stage_advect_density(DT)
        s.sync()
        out = s.download()[0]
    assert_bitwise(out[1:-1, 1:-1], d[1:-1, 1:-1], "interior")
    assert_bitwise(out[-1, 1:-1], d[-2, 1:-1], "bottom edge = last interior row")
