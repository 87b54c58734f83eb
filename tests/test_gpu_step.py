"""GPU parity of the full solve() step through the C ABI: against the oracle (SFO_SEM_GPU), against
the committed reference-GPU golden fixtures, and -- when oracle/_ref/libref_gpu.so travelled to the
box -- against the UNMODIFIED fluid_solver_gpu run live on the same inputs."""
import glob
import os

import numpy as np
import pytest

from util import DIFFUSION_RATE, DT, VISCOSITY, assert_bitwise, assert_close, rng_fields

pytestmark = pytest.mark.gpu
NAIVE, STREAM = 0, 1
DIV_F64, DIV_F32 = 0, 1
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")

# stated tolerances for the density field (downstream of the atomic scatter) per step
D_REL_L2, D_MAX_ABS = 2e-6, 2e-5


def run_steps(f2d, fields, kd, kp, steps, rate=DIFFUSION_RATE, visc=VISCOSITY, **kw):
    d, u, v, sd, su, sv = fields
    with f2d.FluidSolverB200(d.shape[0], d.shape[1], diffuse_iters=kd, project_iters=kp, **kw) as s:
        s.upload(d, u, v)
        s.set_sources(sd, su, sv)
        s.step(rate, visc, DT, steps)
        s.sync()
        return s.download()


@pytest.mark.parametrize("n,kd,kp,steps", [(64, 15, 20, 3), (128, 20, 20, 2), (256, 40, 40, 2), (100, 7, 9, 3)])
@pytest.mark.parametrize("mode,T", [(NAIVE, 1), (STREAM, 4), (STREAM, 8)], ids=["naive", "stream4", "stream8"])
def test_step_vs_oracle(f2d, sfo, gpu_ok, n, kd, kp, steps, mode, T):
    f = rng_fields(n, 900 + n)
    gd, gu, gv = run_steps(f2d, f, kd, kp, steps, jacobi_mode=mode, temporal_block=T, divide_mode=DIV_F64)
    od, ou, ov = sfo.steps(f[0], f[3], DIFFUSION_RATE, f[1], f[2], f[4], f[5], VISCOSITY, DT, kd, kp, nsteps=steps)
    assert_bitwise(gu, ou, "u")  # the velocity chain never sees the atomics
    assert_bitwise(gv, ov, "v")
    assert_close(gd, od, "d", rel_l2=D_REL_L2 * steps, max_abs_rel=D_MAX_ABS * steps)


def test_default_config_within_fp32_tolerance(f2d, sfo, gpu_ok):
    """Product defaults (stream, T=8, fp32-corrected divide, graph) on the canonical C1-like case:
    256^2, (Kd,Kp)=(20,20), 10 free-running steps; tolerance rel-L2 <= 1e-5, max-abs <= 1e-4*|f|max."""
    f = sfo.canonical_fields(256)
    gd, gu, gv = run_steps(f2d, f, 20, 20, 10)
    od, ou, ov = sfo.steps(f[0], f[3], DIFFUSION_RATE, f[1], f[2], f[4], f[5], VISCOSITY, DT, 20, 20, nsteps=10)
    for name, a, b in (("d", gd, od), ("u", gu, ou), ("v", gv, ov)):
        assert_close(a, b, name, rel_l2=1e-5, max_abs_rel=1e-4)


def test_solve_host_is_upload_step_download(f2d, sfo, gpu_ok):
    """fluid_solver::solve semantics (src/fluid_solver.hpp:16-24): in place on host grids, sources const."""
    n = 96
    d, u, v, sd, su, sv = rng_fields(n, 42)
    want = sfo.steps(d, sd, DIFFUSION_RATE, u, v, su, sv, VISCOSITY, DT, 15, 20, nsteps=1)
    hd, hu, hv = d.copy(), u.copy(), v.copy()
    src_before = [a.copy() for a in (sd, su, sv)]
    with f2d.FluidSolverB200(n, n, divide_mode=DIV_F64) as s:
        s.solve(hd, sd, DIFFUSION_RATE, hu, hv, su, sv, VISCOSITY, DT)
        assert_bitwise(hu, want[1], "u")
        assert_bitwise(hv, want[2], "v")
        assert_close(hd, want[0], "d", rel_l2=D_REL_L2, max_abs_rel=D_MAX_ABS)
        for a, b in zip((sd, su, sv), src_before):
            assert np.array_equal(a, b)
        # second call continues from the host state, like the reference
        s.solve(hd, sd, DIFFUSION_RATE, hu, hv, su, sv, VISCOSITY, DT)
        want2 = sfo.steps(want[0], sd, DIFFUSION_RATE, want[1], want[2], su, sv, VISCOSITY, DT, 15, 20, nsteps=1)
        assert_bitwise(hu, want2[1], "u step 2")
        assert s.launch_count() > 0


@pytest.mark.parametrize("graph", [True, False], ids=["graph", "eager"])
@pytest.mark.parametrize("n,kd,kp", [(512, 15, 20), (640, 0, 8), (96, 5, 0)])
def test_pipelined_solve_host_equals_plain_and_keeps_state(f2d, sfo, gpu_ok, monkeypatch, graph, n, kd, kp):
    """solve() overlaps uploads, the four parts of the step and downloads (f2d_solve_host, single GPU).  Same
    bits in u, v as the plain upload -> step -> download path (F2D_HOST_PIPELINE=0) and as the oracle; the
    device-resident state afterwards equals the host grids, so device-resident stepping can continue from it.
    512^2 and 640^2 grids are >= 1 MiB and take the page-locking path (cudaHostRegister)."""
    f = rng_fields(n, 4300 + n)
    want = sfo.steps(f[0], f[3], DIFFUSION_RATE, f[1], f[2], f[4], f[5], VISCOSITY, DT, kd, kp, nsteps=2)
    results = {}
    for pipe in ("1", "0"):
        monkeypatch.setenv("F2D_HOST_PIPELINE", pipe)
        monkeypatch.setenv("F2D_HOST_PIPELINE_MIN_BYTES", "0")  # by default fields under 1 MiB take the serial order
        hd, hu, hv = f[0].copy(), f[1].copy(), f[2].copy()
        with f2d.FluidSolverB200(n, n, diffuse_iters=kd, project_iters=kp, divide_mode=DIV_F64, use_graph=graph) as s:
            for _ in range(2):
                s.solve(hd, f[3], DIFFUSION_RATE, hu, hv, f[4], f[5], VISCOSITY, DT)
            sd_, su_, sv_ = s.download()
            assert_bitwise(su_, hu, "device state u == host u (pipeline=%s)" % pipe)
            assert_bitwise(sv_, hv, "device state v == host v (pipeline=%s)" % pipe)
            assert_bitwise(sd_, hd, "device state d == host d (pipeline=%s)" % pipe)
            # and the device-resident extension continues from there
            s.set_sources(f[3], f[4], f[5])
            s.step(DIFFUSION_RATE, VISCOSITY, DT, 1)
            s.sync()
            results[pipe] = (hd, hu, hv) + s.download()
    for k in (1, 2, 4, 5):
        assert_bitwise(results["1"][k], results["0"][k], "pipelined vs plain, field %d" % k)
    assert_bitwise(results["1"][1], want[1], "u vs oracle")
    assert_bitwise(results["1"][2], want[2], "v vs oracle")
    assert_close(results["1"][0], want[0], "d vs oracle", rel_l2=2 * D_REL_L2, max_abs_rel=2 * D_MAX_ABS)


def test_graph_replay_equals_eager(f2d, gpu_ok):
    n = 192
    f = rng_fields(n, 43)
    a = run_steps(f2d, f, 15, 20, 4, use_graph=True, divide_mode=DIV_F64)
    b = run_steps(f2d, f, 15, 20, 4, use_graph=False, divide_mode=DIV_F64)
    assert_bitwise(a[1], b[1], "u")
    assert_bitwise(a[2], b[2], "v")
    assert_close(a[0], b[0], "d", rel_l2=4 * D_REL_L2, max_abs_rel=4 * D_MAX_ABS)


def test_nonsquare_and_odd_sizes_naive_vs_oracle(f2d, sfo, gpu_ok):
    """The reference itself is square-only (grid<T>::cols() returns rows); the oracle and the CUDA path
    agree on the natural generalisation.  Odd column counts take the naive relaxation."""
    for rows, cols in ((37, 37), (50, 50), (48, 80), (33, 20)):
        f = rng_fields(rows, 1000 + rows, cols=cols)
        mode = STREAM if cols % 4 == 0 else NAIVE
        g = run_steps(f2d, f, 6, 9, 2, jacobi_mode=mode, temporal_block=4 if mode == STREAM else 1, divide_mode=DIV_F64)
        o = sfo.steps(f[0], f[3], DIFFUSION_RATE, f[1], f[2], f[4], f[5], VISCOSITY, DT, 6, 9, nsteps=2)
        assert_bitwise(g[1], o[1], "u %dx%d" % (rows, cols))
        assert_bitwise(g[2], o[2], "v %dx%d" % (rows, cols))
        assert_close(g[0], o[0], "d %dx%d" % (rows, cols), rel_l2=2 * D_REL_L2, max_abs_rel=2 * D_MAX_ABS)


def test_stream_rejects_unsupported_geometry(f2d, gpu_ok):
    with pytest.raises(f2d.F2DError):
        f2d.FluidSolverB200(37, 37, jacobi_mode=STREAM)


FIXTURES = sorted(glob.glob(os.path.join(GOLDEN, "refgpu_*.npz")))


@pytest.mark.skipif(not FIXTURES, reason="no refgpu fixtures committed yet")
@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p) for p in FIXTURES])
def test_cuda_path_matches_reference_gpu_fixture(f2d, gpu_ok, path):
    """CUDA path vs golden vectors of the unmodified fluid_solver_gpu::solve (Kd=15, Kp=20, smooth)."""
    z = np.load(path)
    f = tuple(z[k] for k in ("d", "u", "v", "sd", "su", "sv"))
    for steps, tag in ((1, "solve1"), (3, "solve3")):
        gd, gu, gv = run_steps(f2d, f, 15, 20, steps, rate=float(z["diffusion_rate"]), visc=float(z["viscosity"]),
                               divide_mode=DIV_F64)
        assert_bitwise(gu, z[tag + "_u"], tag + " u")
        assert_bitwise(gv, z[tag + "_v"], tag + " v")
        assert_close(gd, z[tag + "_d"], tag + " d", rel_l2=D_REL_L2 * steps, max_abs_rel=D_MAX_ABS * steps)


def test_live_against_unmodified_reference_gpu_solver(f2d, gpu_ok):
    from oracle import refs

    if not refs.have_gpu():
        pytest.skip("oracle/_ref/libref_gpu.so did not travel to this box")
    g = refs.ref_gpu()
    for n, seed in ((64, 1), (256, 2), (512, 3)):
        f = rng_fields(n, seed)
        rd, ru, rv, _ = g.solve(f[0], f[3], DIFFUSION_RATE, f[1], f[2], f[4], f[5], VISCOSITY, DT, 2)
        gd, gu, gv = run_steps(f2d, f, 15, 20, 2, divide_mode=DIV_F64)
        assert_bitwise(gu, ru, "u n=%d" % n)
        assert_bitwise(gv, rv, "v n=%d" % n)
        assert_close(gd, rd, "d n=%d" % n, rel_l2=2 * D_REL_L2, max_abs_rel=2 * D_MAX_ABS)
        # product defaults (fp32-corrected divide): within the stated fp32 tolerance
        pd, pu, pv = run_steps(f2d, f, 15, 20, 2)
        for name, a, b in (("d", pd, rd), ("u", pu, ru), ("v", pv, rv)):
            assert_close(a, b, "default %s n=%d" % (name, n), rel_l2=4e-6, max_abs_rel=4e-5)


def test_headless_renderers_match_restatement(f2d, gpu_ok):
    """density -> RGBA8 (src/density_grid_renderer.cu:10-29) and velocity -> line list
    (src/velocity_grid_renderer.cu:8-44) on the device-resident fields, against a numpy restatement
    (a second, independent check; the reference's own renderers are the pin: tests/test_gpu_render_ref.py)."""
    n = 72
    d, u, v, *_ = rng_fields(n, 77)
    d = (d * np.float32(3.0) - np.float32(0.5)).astype(np.float32)  # exercise both clamps
    with f2d.FluidSolverB200(n, n) as s:
        s.upload(d, u, v)
        img = s.render_density_rgba((255.0, 160.0, 64.0))
        ln = s.render_velocity_lines(10.0, 7.5)
    for c, m in enumerate((255.0, 160.0, 64.0)):
        want = np.clip(np.float32(m) * d, np.float32(0), np.float32(255)).astype(np.uint8)
        assert np.array_equal(img[:, :, c], want)
    assert (img[:, :, 3] == 255).all()
    jj, ii = np.meshgrid(np.arange(n, dtype=np.float32), np.arange(n, dtype=np.float32))
    sx, sy = jj * np.float32(10.0), ii * np.float32(7.5)
    ex, ey = sx.copy(), sy.copy()
    norm = np.sqrt(np.float32(n * n))
    sel = (np.arange(n) % 8 == 0)
    m = np.outer(sel, sel)
    ex[m] = ex[m] + (np.float32(250000.0) * u[m]) / norm
    ey[m] = ey[m] + (np.float32(250000.0) * v[m]) / norm
    assert np.array_equal(ln[:, :, 0], sx) and np.array_equal(ln[:, :, 1], sy)
    assert_bitwise(np.ascontiguousarray(ln[:, :, 2]), ex, "line end x")
    assert_bitwise(np.ascontiguousarray(ln[:, :, 3]), ey, "line end y")


def test_full_step_at_the_headline_config_vs_live_reference_gpu(f2d, gpu_ok):
    """BASELINE configs[2], the configuration bench.py's headline is quoted on: 4096^2, Kd = Kp = 80, canonical
    fields, one full step against the UNMODIFIED fluid_solver_gpu run live with the same iteration counts
    (stage sequence of src/fluid_solver_gpu.cu:236-252 through oracle/ref_gpu_shim.cu).
      * F2D_DIV_F64: u and v bit-identical; density within the atomic-order tolerance (rel-L2 <= 2e-6,
        max-abs <= 2e-5 * max(1, |d|max));
      * product defaults (fp32-corrected divide, a = 1.7e5 here): every field within rel-L2 <= 2e-6,
        max-abs <= 2e-5 * max(1, |f|max) per step."""
    from oracle import refs
    from tools import canonical

    if not refs.have_gpu():
        pytest.skip("oracle/_ref/libref_gpu.so did not travel to this box")
    n, k = 4096, 80
    f = canonical.fields(n)
    rd, ru, rv, _ = refs.ref_gpu().step_k(f[0], f[3], DIFFUSION_RATE, f[1], f[2], f[4], f[5], VISCOSITY, DT, k, k, True, 1)
    gd, gu, gv = run_steps(f2d, f, k, k, 1, divide_mode=DIV_F64)
    assert_bitwise(gu, ru, "u (fp64 divide)")
    assert_bitwise(gv, rv, "v (fp64 divide)")
    assert_close(gd, rd, "d (fp64 divide)", rel_l2=D_REL_L2, max_abs_rel=D_MAX_ABS)
    pd, pu, pv = run_steps(f2d, f, k, k, 1)
    for name, a, b in (("d", pd, rd), ("u", pu, ru), ("v", pv, rv)):
        e = assert_close(a, b, "default %s" % name, rel_l2=2e-6, max_abs_rel=2e-5)
        print("4096^2 K=80 default config vs fluid_solver_gpu: %s %r" % (name, e))
