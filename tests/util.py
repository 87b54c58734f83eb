"""Shared helpers for the parity tests: seeded inputs and error metrics."""
import numpy as np

DT = 0.02
DIFFUSION_RATE = 0.5   # src/app.cpp:33
VISCOSITY = 1e-6       # src/app.cpp:34


def rng_fields(n, seed, cols=None, vel_cells=3.0, dt=DT):
    """Seeded random-but-smooth test fields (d,u,v,sd,su,sv) on an n x cols grid.  Velocities are
    scaled so that the largest displacement per step is about `vel_cells` cells
    (displacement = dt * N * |v|, N = sqrt(rows*cols))."""
    cols = cols or n
    r = np.random.default_rng(seed)
    y, x = np.meshgrid((np.arange(n) + 0.5) / n, (np.arange(cols) + 0.5) / cols, indexing="ij")
    amp = vel_cells / (np.sqrt(n * cols) * dt)

    def smooth_noise(scale):
        a = np.zeros((n, cols))
        for _ in range(4):
            kx, ky = r.integers(1, 5, size=2)
            ph = r.uniform(0, 2 * np.pi, size=2)
            a += r.normal() * np.sin(2 * np.pi * kx * x + ph[0]) * np.cos(2 * np.pi * ky * y + ph[1])
        a += 0.1 * r.normal(size=(n, cols))
        return (scale * a / max(1e-12, np.abs(a).max())).astype(np.float32)

    d = np.abs(smooth_noise(1.0)) + np.float32(0.01)
    u = smooth_noise(amp)
    v = smooth_noise(amp)
    sd = (r.uniform(size=(n, cols)) < 0.02).astype(np.float32) * np.float32(1.0)
    su = smooth_noise(amp) * (r.uniform(size=(n, cols)) < 0.05)
    sv = smooth_noise(amp) * (r.uniform(size=(n, cols)) < 0.05)
    return tuple(np.ascontiguousarray(a, dtype=np.float32) for a in (d, u, v, sd, su, sv))


def ulp_diff(a, b):
    """Element-wise distance in units in the last place (fp32), sign-magnitude aware."""
    ai = a.view(np.int32).astype(np.int64)
    bi = b.view(np.int32).astype(np.int64)
    ai = np.where(ai < 0, -(ai & 0x7FFFFFFF), ai)
    bi = np.where(bi < 0, -(bi & 0x7FFFFFFF), bi)
    return np.abs(ai - bi)


def err(a, ref):
    """dict(max_abs, rel_l2, max_ulp, n_diff) of a against ref."""
    a64, r64 = a.astype(np.float64), ref.astype(np.float64)
    d = a64 - r64
    nrm = np.sqrt((r64 * r64).sum())
    u = ulp_diff(a, ref)
    return dict(max_abs=float(np.abs(d).max()), rel_l2=float(np.sqrt((d * d).sum()) / max(nrm, 1e-300)),
                max_ulp=int(u.max()), n_diff=int((a.view(np.int32) != ref.view(np.int32)).sum()))


def assert_bitwise(a, ref, what):
    if not np.array_equal(a.view(np.int32), ref.view(np.int32)):
        e = err(a, ref)
        bad = np.argwhere(a.view(np.int32) != ref.view(np.int32))
        i, j = bad[0]
        raise AssertionError("%s not bit-identical: %r; first mismatch at (%d,%d): got %r want %r; %d cells differ, "
                             "rows %d..%d cols %d..%d" % (what, e, i, j, a[i, j], ref[i, j], len(bad),
                                                          bad[:, 0].min(), bad[:, 0].max(), bad[:, 1].min(), bad[:, 1].max()))


def assert_close(a, ref, what, max_ulp=None, rel_l2=None, max_abs_rel=None, max_frac=None):
    """Tolerance check with the tolerance written at the call site."""
    e = err(a, ref)
    scale = max(1.0, float(np.abs(ref).max()))
    msg = "%s: %r" % (what, e)
    if max_ulp is not None:
        assert e["max_ulp"] <= max_ulp, msg
    if rel_l2 is not None:
        assert e["rel_l2"] <= rel_l2, msg
    if max_abs_rel is not None:
        assert e["max_abs"] <= max_abs_rel * scale, msg
    if max_frac is not None:
        assert e["n_diff"] <= max_frac * a.size, msg
    return e
