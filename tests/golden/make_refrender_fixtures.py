"""Generates the golden vectors that pin the headless renderers.

Runs the UNMODIFIED reference renderers (oracle/_ref/libref_render.so = /root/reference/src/
density_grid_renderer.cu + velocity_grid_renderer.cu compiled with nvcc defaults for sm_100a against
oracle/sfml_stub + oracle/ref_render_shim.cu) on seeded fields and stores inputs and outputs.  Needs a GPU:

    gpurun -- 'python tests/golden/make_refrender_fixtures.py gpurun_out/golden'

then copy gpurun_out/golden/refrender_*.npz into tests/golden/ and commit them.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

from util import rng_fields  # noqa: E402

CASES = [(72, 501, (800, 800), (255.0, 160.0, 64.0)), (40, 502, (640, 480), (300.0, 255.0, 10.0))]


def main(outdir):
    from oracle import refs

    assert refs.have_render(), "needs oracle/_ref/libref_render.so and a CUDA device"
    os.makedirs(outdir, exist_ok=True)
    for n, seed, target, mult in CASES:
        d, u, v, *_ = rng_fields(n, seed)
        d = (d * np.float32(3.0) - np.float32(0.5)).astype(np.float32)  # exercise both clamps
        img = refs.ref_render_density(d, mult, target)
        lines = refs.ref_render_velocity(u, v, target)
        np.savez_compressed(os.path.join(outdir, "refrender_%d.npz" % n), n=n, seed=seed, target=np.array(target), mult=np.array(mult, np.float32),
                            d=d, u=u, v=v, img=img, lines=lines)
        print("refrender_%d.npz: image sum %d, line checksum %.6e" % (n, int(img.astype(np.int64).sum()), float(np.abs(lines).sum())))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
