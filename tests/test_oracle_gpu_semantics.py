"""SFO_SEM_GPU (the parity target) checked on the CPU:
 * against an independent vectorised numpy restatement of the same contract
   where numpy can express the arithmetic exactly (no FMA involved: pressure solve, smooth,
   set_bnd, divergence/gradient with power-of-two N);
 * through properties: boundary invariants, corners untouched, Jacobi fixed point, scatter mass;
 * against the golden fixtures produced by the UNMODIFIED fluid_solver_gpu on a B200
   (tests/golden/refgpu_*.npz, generator tests/golden/make_refgpu_fixtures.py)."""
import glob
import os

import numpy as np
import pytest

from util import DIFFUSION_RATE, DT, VISCOSITY, assert_bitwise, assert_close, rng_fields

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def np_set_bnd_gpu(f, kind):
    f = f.copy()
    sc = -1.0 if kind == 1 else 1.0
    sr = -1.0 if kind == 2 else 1.0
    f[1:-1, 0] = np.float32(sc) * f[1:-1, 1]
    f[1:-1, -1] = np.float32(sc) * f[1:-1, -2]
    f[0, 1:-1] = np.float32(sr) * f[1, 1:-1]
    f[-1, 1:-1] = np.float32(sr) * f[-2, 1:-1]
    return f


def np_project_gpu(u, v, iters):
    n = u.shape[0]
    h = np.float32(1.0) / np.sqrt(np.float32(n * n))
    dv = np.zeros_like(u)
    s = (u[1:-1, 2:] - u[1:-1, :-2]) + v[2:, 1:-1]
    s = s - v[:-2, 1:-1]
    dv[1:-1, 1:-1] = (np.float32(-0.5) * h) * s
    dv = np_set_bnd_gpu(dv, 0)
    p = np.zeros_like(u)
    for _ in range(iters):
        pp = p.copy()
        s = dv[1:-1, 1:-1] + pp[1:-1, 2:]
        s = s + pp[1:-1, :-2]
        s = s + pp[2:, 1:-1]
        s = s + pp[:-2, 1:-1]
        p[1:-1, 1:-1] = s * np.float32(0.25)
        p = np_set_bnd_gpu(p, 0)
    u, v = u.copy(), v.copy()
    u[1:-1, 1:-1] = u[1:-1, 1:-1] - (np.float32(0.5) * (p[1:-1, 2:] - p[1:-1, :-2])) / h
    v[1:-1, 1:-1] = v[1:-1, 1:-1] - (np.float32(0.5) * (p[2:, 1:-1] - p[:-2, 1:-1])) / h
    return np_set_bnd_gpu(u, 1), np_set_bnd_gpu(v, 2), p, dv


@pytest.mark.parametrize("n", [8, 17, 64, 96])
def test_project_matches_numpy_restatement(sfo, n):
    d, u, v, *_ = rng_fields(n, 7 + n)
    ou, ov, op, odv = sfo.project(u, v, 13, sfo.SEM_GPU, return_p=True)
    nu, nv, npp, ndv = np_project_gpu(u, v, 13)
    assert_bitwise(odv, ndv, "divergence")
    assert_bitwise(op, npp, "pressure")
    assert_bitwise(ou, nu, "u")
    assert_bitwise(ov, nv, "v")


@pytest.mark.parametrize("kind", [0, 1, 2])
def test_set_bnd_gpu_leaves_corners(sfo, kind):
    d, *_ = rng_fields(21, 3)
    o = sfo.set_bnd(d, kind, sfo.SEM_GPU)
    assert_bitwise(o, np_set_bnd_gpu(d, kind), "set_bnd")
    for c in ((0, 0), (0, -1), (-1, 0), (-1, -1)):
        assert o[c] == d[c]


def test_smooth_matches_numpy(sfo):
    d, *_ = rng_fields(40, 5)
    o = sfo.smooth(d)
    s = d[1:-1, 1:-1] + d[1:-1, :-2]
    s = s + d[1:-1, 2:]
    s = s + d[:-2, 1:-1]
    s = s + d[2:, 1:-1]
    want = d.copy()
    want[1:-1, 1:-1] = np.float32(0.2) * s
    assert_bitwise(o, want, "smooth")


def test_diffuse_jacobi_fixed_point(sfo):
    """Many Jacobi sweeps converge to the solution of (1+4a) x - a*sum4(x) = x0 on the interior."""
    n = 24
    d, *_ = rng_fields(n, 11)
    rate = 2e-3
    o = sfo.diffuse(d, 0, rate, DT, 4000, sfo.SEM_GPU)
    a = float(sfo.lib().sfo_diffuse_coeff(n, n, rate, DT))
    o64 = o.astype(np.float64)
    res = (1 + 4 * a) * o64[1:-1, 1:-1] - a * (o64[1:-1, :-2] + o64[1:-1, 2:] + o64[:-2, 1:-1] + o64[2:, 1:-1]) - d[1:-1, 1:-1]
    assert np.abs(res).max() < 1e-5 * max(1.0, 1 + 4 * a)


def test_diffuse_jacobi_differs_from_gauss_seidel(sfo):
    """The two reference solvers are not equivalent: guards against the oracle
    silently using the wrong relaxation."""
    d, *_ = rng_fields(32, 12)
    gj = sfo.diffuse(d, 0, DIFFUSION_RATE, DT, 20, sfo.SEM_GPU)
    gs = sfo.diffuse(d, 0, DIFFUSION_RATE, DT, 20, sfo.SEM_CPU)
    assert np.abs(gj - gs).max() > 1e-4


def test_scatter_conserves_mass_for_interior_flow(sfo):
    n = 64
    d, u, v, *_ = rng_fields(n, 13, vel_cells=2.0)
    d[:8, :] = 0
    d[-8:, :] = 0
    d[:, :8] = 0
    d[:, -8:] = 0
    o = sfo.advect_scatter(d, u, v, 0, DT, sfo.SEM_GPU)
    # mass moves by <= 2 cells: nothing reaches the edges, so the interior sums agree
    assert abs(o[1:-1, 1:-1].astype(np.float64).sum() - d.astype(np.float64).sum()) < 1e-3


def test_gather_of_constant_field_is_constant(sfo):
    n = 48
    _, u, v, *_ = rng_fields(n, 14, vel_cells=5.0)
    c = np.full((n, n), 3.25, np.float32)
    o = sfo.advect_gather(c, u, v, 0, DT, sfo.SEM_GPU)
    assert np.abs(o[1:-1, 1:-1] - 3.25).max() <= 5e-7 * 3.25


def test_gpu_and_cpu_semantics_agree_where_algorithms_coincide(sfo):
    """add_sources / advect differ only by FMA contraction: a few ulp."""
    d, u, v, sd, su, sv = rng_fields(64, 15)
    assert_close(sfo.add_sources(d, sd, DT, sfo.SEM_GPU), sfo.add_sources(d, sd, DT, sfo.SEM_CPU), "add", max_ulp=2)
    g, c = sfo.advect_gather(d, u, v, 0, DT, sfo.SEM_GPU), sfo.advect_gather(d, u, v, 0, DT, sfo.SEM_CPU)
    for corner in ((0, 0), (0, -1), (-1, 0), (-1, -1)):  # GPU: untouched, CPU: averaged
        assert g[corner] == d[corner]
        c[corner] = g[corner]
    assert_close(g, c, "gather", rel_l2=1e-6, max_abs_rel=4e-6)


FIXTURES = sorted(glob.glob(os.path.join(GOLDEN, "refgpu_*.npz")))


@pytest.mark.skipif(not FIXTURES, reason="no refgpu fixtures committed yet (parity unpinned against the GPU solver)")
@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p) for p in FIXTURES])
def test_oracle_matches_reference_gpu_fixture(sfo, path):
    """Golden vectors written by the UNMODIFIED fluid_solver_gpu on a B200 pin SFO_SEM_GPU."""
    from golden.check_refgpu_fixture import check_fixture

    check_fixture(sfo, np.load(path))
