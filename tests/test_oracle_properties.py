"""Property tests (hypothesis) of the oracle's GPU semantics: the size-independent invariants the
GPU parity tests rely on at full size."""
import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

from util import DT, rng_fields

sizes = st.sampled_from([5, 8, 13, 24, 40])
seeds = st.integers(min_value=0, max_value=10_000)


@settings(max_examples=25, deadline=None)
@given(n=sizes, seed=seeds, kind=st.sampled_from([0, 1, 2]))
def test_set_bnd_invariants(sfo, n, seed, kind):
    """Edges are +/- the adjacent interior cell, corners and interior are untouched, and the pass is idempotent."""
    f = rng_fields(n, seed)[1]
    o = sfo.set_bnd(f, kind)
    sc = -1.0 if kind == 1 else 1.0
    sr = -1.0 if kind == 2 else 1.0
    assert np.array_equal(o[1:-1, 1:-1], f[1:-1, 1:-1])
    assert np.array_equal(o[1:-1, 0], np.float32(sc) * f[1:-1, 1]) and np.array_equal(o[1:-1, -1], np.float32(sc) * f[1:-1, -2])
    assert np.array_equal(o[0, 1:-1], np.float32(sr) * f[1, 1:-1]) and np.array_equal(o[-1, 1:-1], np.float32(sr) * f[-2, 1:-1])
    for c in ((0, 0), (0, -1), (-1, 0), (-1, -1)):
        assert o[c] == f[c]
    assert np.array_equal(sfo.set_bnd(o, kind), o)


@settings(max_examples=15, deadline=None)
@given(n=sizes, seed=seeds, k1=st.integers(0, 6), k2=st.integers(0, 6))
def test_pressure_sweeps_compose(sfo, n, seed, k1, k2):
    """K1 + K2 Jacobi sweeps of the pressure solve == K1 sweeps continued by K2 more (what temporal blocking
    relies on): checked through the returned pressure field with an independent continuation in numpy."""
    _, u, v, *_ = rng_fields(n, seed)
    _, _, p_all, dv = sfo.project(u, v, k1 + k2, return_p=True)
    _, _, p1, dv1 = sfo.project(u, v, k1, return_p=True)
    assert np.array_equal(dv, dv1)
    p = p1.copy()
    for _ in range(k2):
        pp = p.copy()
        s = dv[1:-1, 1:-1] + pp[1:-1, 2:]
        s = s + pp[1:-1, :-2]
        s = s + pp[2:, 1:-1]
        s = s + pp[:-2, 1:-1]
        p[1:-1, 1:-1] = s * np.float32(0.25)
        p = sfo.set_bnd(p, 0)
    assert np.array_equal(p, p_all)


@settings(max_examples=15, deadline=None)
@given(n=sizes, seed=seeds)
def test_first_pressure_sweep_is_quarter_divergence(sfo, n, seed):
    """p0 == 0, so sweep 1 is exactly 0.25*div on the interior: the licence for fusing the divergence into
    the first pressure pass."""
    _, u, v, *_ = rng_fields(n, seed)
    _, _, p1, dv = sfo.project(u, v, 1, return_p=True)
    assert np.array_equal(p1[1:-1, 1:-1], dv[1:-1, 1:-1] * np.float32(0.25))


@settings(max_examples=15, deadline=None)
@given(n=sizes, seed=seeds, cells=st.floats(0.1, 3.0))
def test_scatter_is_linear_and_mass_conserving_inside(sfo, n, seed, cells):
    d, u, v, *_ = rng_fields(max(n, 16), seed, vel_cells=cells)
    m = d.shape[0]
    d[:5] = 0
    d[-5:] = 0
    d[:, :5] = 0
    d[:, -5:] = 0
    a = sfo.advect_scatter(d, u, v, 0, DT)
    b = sfo.advect_scatter(np.float32(2.0) * d, u, v, 0, DT)
    assert np.array_equal(b, np.float32(2.0) * a)  # scaling by a power of two is exact
    if m >= 16:
        assert abs(float(a[1:-1, 1:-1].astype(np.float64).sum()) - float(d.astype(np.float64).sum())) <= 1e-4 * max(1.0, float(d.sum()))


@settings(max_examples=15, deadline=None)
@given(n=sizes, seed=seeds)
def test_zero_velocity_advection_is_identity_on_interior(sfo, n, seed):
    d = rng_fields(n, seed)[0]
    z = np.zeros_like(d)
    g = sfo.advect_gather(d, z, z, 0, DT)
    assert np.array_equal(g[2:-2, 2:-2], d[2:-2, 2:-2])  # cells next to the edge are clamped to 1.5 / N-1.5
    s = sfo.advect_scatter(d, z, z, 0, DT)
    assert np.array_equal(s[1:-1, 1:-1], d[1:-1, 1:-1])


@settings(max_examples=10, deadline=None)
@given(n=sizes, seed=seeds, kd=st.integers(0, 5), kp=st.integers(0, 5))
def test_step_is_deterministic_and_leaves_inputs_alone(sfo, n, seed, kd, kp):
    f = rng_fields(n, seed)
    keep = [a.copy() for a in f]
    r1 = sfo.steps(f[0], f[3], 0.5, f[1], f[2], f[4], f[5], 1e-6, DT, kd, kp, nsteps=2)
    r2 = sfo.steps(f[0], f[3], 0.5, f[1], f[2], f[4], f[5], 1e-6, DT, kd, kp, nsteps=2)
    assert all(np.array_equal(a, b) for a, b in zip(r1, r2))
    assert all(np.array_equal(a, b) for a, b in zip(f, keep))
    assert all(np.isfinite(a).all() for a in r1)
