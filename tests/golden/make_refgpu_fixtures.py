"""Generates the golden vectors that pin the oracle's GPU semantics (SFO_SEM_GPU).

Runs the UNMODIFIED reference GPU solver (oracle/_ref/libref_gpu.so = /root/reference/src/
fluid_solver_gpu.cu compiled with nvcc defaults for sm_100a + oracle/ref_gpu_shim.cu) on seeded
inputs and stores inputs and outputs.  Must run on a box with a GPU:

    gpurun -- 'python tests/golden/make_refgpu_fixtures.py gpurun_out/golden'

then copy gpurun_out/golden/refgpu_*.npz into tests/golden/ and commit them.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

from util import DIFFUSION_RATE, DT, VISCOSITY, rng_fields  # noqa: E402

CASES = [(32, 201), (64, 202), (96, 203)]


def main(outdir):
    from oracle import refs

    assert refs.have_gpu(), "needs oracle/_ref/libref_gpu.so and a CUDA device"
    g = refs.ref_gpu()
    os.makedirs(outdir, exist_ok=True)
    for n, seed in CASES:
        d, u, v, sd, su, sv = rng_fields(n, seed, vel_cells=4.0)
        out = dict(n=n, seed=seed, dt=DT, diffusion_rate=DIFFUSION_RATE, viscosity=VISCOSITY,
                   d=d, u=u, v=v, sd=sd, su=su, sv=sv)
        for kind in (0, 1, 2):
            out["set_bnd_%d" % kind] = g.set_bnd(u, kind)
            out["diffuse_%d_hi" % kind] = g.diffuse(d, kind, DIFFUSION_RATE, DT, 15)
            out["diffuse_%d_lo" % kind] = g.diffuse(u, kind, VISCOSITY, DT, 15)
            out["diffuse_%d_mid" % kind] = g.diffuse(u, kind, 1e-4, DT, 6)
            out["gather_%d" % kind] = g.advect(d, u, v, kind, DT, False)
        out["add_sources"] = g.add_sources(d, sd, DT)
        out["smooth"] = g.smooth(d)
        out["scatter"] = g.advect(d, u, v, 0, DT, True)
        out["scatter_again"] = g.advect(d, u, v, 0, DT, True)  # run-to-run atomic-order noise
        pu, pv, pp, pdv = g.project(u, v, 20, return_p=True)
        out.update(project_u=pu, project_v=pv, project_p=pp, project_div=pdv)
        # the reference's own solve(): Kd=15, Kp=20, smooth on (src/fluid_solver_gpu.cu:236-252)
        sd1, su1, sv1, _ = g.solve(d, sd, DIFFUSION_RATE, u, v, su, sv, VISCOSITY, DT, 1)
        out.update(solve1_d=sd1, solve1_u=su1, solve1_v=sv1)
        sd3, su3, sv3, _ = g.solve(d, sd, DIFFUSION_RATE, u, v, su, sv, VISCOSITY, DT, 3)
        out.update(solve3_d=sd3, solve3_u=su3, solve3_v=sv3)
        # free iteration counts through the private stage methods
        kd, kp = 7, 9
        d2, u2, v2, _ = g.step_k(d, sd, 1e-4, u, v, su, sv, 1e-4, DT, kd, kp, smooth=True, nsteps=2)
        out.update(stepk_kd=kd, stepk_kp=kp, stepk_d=d2, stepk_u=u2, stepk_v=v2)
        path = os.path.join(outdir, "refgpu_%d.npz" % n)
        np.savez_compressed(path, **out)
        print("wrote", path)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
