"""Checks the oracle (SFO_SEM_GPU) against one golden fixture written by the unmodified
fluid_solver_gpu (see make_refgpu_fixtures.py).  Deterministic stages must be BIT-IDENTICAL; the
density scatter uses float atomics in hardware order in the reference (src/fluid_solver_gpu.cu:156-159),
so anything downstream of it is compared with a tolerance calibrated by the reference's own
run-to-run spread (`scatter` vs `scatter_again`)."""
import numpy as np

from util import assert_bitwise, assert_close, err


def check_fixture(sfo, z):
    n = int(z["n"])
    dt, rate, visc = float(z["dt"]), float(z["diffusion_rate"]), float(z["viscosity"])
    d, u, v, sd, su, sv = (z[k] for k in ("d", "u", "v", "sd", "su", "sv"))
    G = sfo.SEM_GPU
    for kind in (0, 1, 2):
        assert_bitwise(sfo.set_bnd(u, kind, G), z["set_bnd_%d" % kind], "set_bnd %d" % kind)
        assert_bitwise(sfo.diffuse(d, kind, rate, dt, 15, G), z["diffuse_%d_hi" % kind], "diffuse hi %d" % kind)
        assert_bitwise(sfo.diffuse(u, kind, visc, dt, 15, G), z["diffuse_%d_lo" % kind], "diffuse lo %d" % kind)
        assert_bitwise(sfo.diffuse(u, kind, 1e-4, dt, 6, G), z["diffuse_%d_mid" % kind], "diffuse mid %d" % kind)
        assert_bitwise(sfo.advect_gather(d, u, v, kind, dt, G), z["gather_%d" % kind], "gather %d" % kind)
    assert_bitwise(sfo.add_sources(d, sd, dt, G), z["add_sources"], "add_sources")
    assert_bitwise(sfo.smooth(d), z["smooth"], "smooth")
    pu, pv, pp, pdv = sfo.project(u, v, 20, G, return_p=True)
    assert_bitwise(pdv, z["project_div"], "divergence")
    assert_bitwise(pp, z["project_p"], "pressure")
    assert_bitwise(pu, z["project_u"], "project u")
    assert_bitwise(pv, z["project_v"], "project v")
    # scatter: tolerance = a few ulp of the field maximum (atomic summation order)
    noise = err(z["scatter_again"], z["scatter"])
    e = assert_close(sfo.advect_scatter(d, u, v, 0, dt, G), z["scatter"], "scatter", max_abs_rel=1e-6, rel_l2=5e-7)
    assert e["max_abs"] <= max(8 * noise["max_abs"], 1e-6 * max(1.0, float(np.abs(z["scatter"]).max())))
    # full steps: u, v never see the atomics -> bitwise; density within the stated tolerance
    for steps, tag in ((1, "solve1"), (3, "solve3")):
        od, ou, ov = sfo.steps(d, sd, rate, u, v, su, sv, visc, dt, 15, 20, smooth=True, sem=G, nsteps=steps)
        assert_bitwise(ou, z[tag + "_u"], tag + " u")
        assert_bitwise(ov, z[tag + "_v"], tag + " v")
        assert_close(od, z[tag + "_d"], tag + " d", rel_l2=2e-6, max_abs_rel=2e-5)
    kd, kp = int(z["stepk_kd"]), int(z["stepk_kp"])
    od, ou, ov = sfo.steps(d, sd, 1e-4, u, v, su, sv, 1e-4, dt, kd, kp, smooth=True, sem=G, nsteps=2)
    assert_bitwise(ou, z["stepk_u"], "stepk u")
    assert_bitwise(ov, z["stepk_v"], "stepk v")
    assert_close(od, z["stepk_d"], "stepk d", rel_l2=2e-6, max_abs_rel=2e-5)
