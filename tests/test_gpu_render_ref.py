"""the headless renderers against the UNMODIFIED reference renderers
(src/density_grid_renderer.cu:10-56, src/velocity_grid_renderer.cu:8-72), which are compiled out of
/root/reference against oracle/sfml_stub (a recording stand-in for the SFML types they touch; oracle/Makefile).
Live when oracle/_ref/libref_render.so travelled to the box, and against golden vectors it produced on a B200
(tests/golden/refrender_*.npz, tests/golden/make_refrender_fixtures.py).  Bit-exact: bytes of the image,
bits of the segment coordinates."""
import glob
import os

import numpy as np
import pytest

from util import assert_bitwise, rng_fields

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def ours(f2d, d, u, v, mult, target):
    n = d.shape[0]
    # the reference forms the scales as float(target) / grid size (src/velocity_grid_renderer.cu:64)
    hs = np.float32(target[0]) / np.float32(n)
    vs = np.float32(target[1]) / np.float32(n)
    with f2d.FluidSolverB200(n, n) as s:
        s.upload(d, u, v)
        return s.render_density_rgba(tuple(float(m) for m in mult)), s.render_velocity_lines(float(hs), float(vs))


def check(img, lines, want_img, want_lines, tag):
    assert np.array_equal(img, want_img), "%s: %d image bytes differ" % (tag, int((img != want_img).sum()))
    for k, name in enumerate(("start.x", "start.y", "end.x", "end.y")):
        assert_bitwise(np.ascontiguousarray(lines[:, :, k]), np.ascontiguousarray(want_lines[:, :, k]), "%s %s" % (tag, name))


@pytest.mark.parametrize("n,target", [(64, (800, 800)), (200, (1024, 768)), (33, (100, 900))])
def test_live_against_unmodified_reference_renderers(f2d, gpu_ok, n, target):
    from oracle import refs

    if not refs.have_render():
        pytest.skip("oracle/_ref/libref_render.so did not travel to this box")
    d, u, v, *_ = rng_fields(n, 600 + n)
    d = (d * np.float32(3.0) - np.float32(0.5)).astype(np.float32)
    mult = (255.0, 160.0, 64.0)
    img, lines = ours(f2d, d, u, v, mult, target)
    check(img, lines, refs.ref_render_density(d, mult, target), refs.ref_render_velocity(u, v, target), "n=%d" % n)


FIXTURES = sorted(glob.glob(os.path.join(GOLDEN, "refrender_*.npz")))


@pytest.mark.skipif(not FIXTURES, reason="no refrender fixtures committed yet")
@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p) for p in FIXTURES])
def test_against_reference_renderer_fixture(f2d, gpu_ok, path):
    z = np.load(path)
    img, lines = ours(f2d, z["d"], z["u"], z["v"], z["mult"], tuple(int(t) for t in z["target"]))
    check(img, lines, z["img"], z["lines"], os.path.basename(path))
