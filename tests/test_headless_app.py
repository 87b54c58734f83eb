"""The headless caller loop (app/fluid2d_headless.cpp, include/simulation_headless.hpp): the reference's
simulation::update contract (src/simulation.cpp:44-65) around the drop-in solver, and the .npy field
exchange format (include/f2d_npy.hpp)."""
import math
import os
import subprocess

import numpy as np
import pytest

from util import assert_bitwise, assert_close

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
APP = os.path.join(ROOT, "app", "fluid2d_headless")


def build_app(f2d):
    subprocess.run(["make", "-C", os.path.join(ROOT, "app")], check=True, stdout=subprocess.DEVNULL)
    return APP


def test_npy_header_roundtrip(tmp_path):
    """f2d_npy::save output is readable by numpy and f2d_npy::load reads numpy's output."""
    src = tmp_path / "t.cpp"
    src.write_text('''#include "f2d_npy.hpp"
#include <cstdio>
int main(int argc, char** argv) {
    std::vector<float> d; size_t r, c;
    f2d_npy::load(argv[1], d, r, c);
    for (auto& x : d) x = 2.0f * x + 1.0f;
    f2d_npy::save(argv[2], d.data(), r, c);
    std::printf("%zu %zu\\n", r, c);
    return 0;
}''')
    exe = tmp_path / "t"
    subprocess.run(["g++", "-std=c++14", "-O1", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    for shape in ((3, 5), (64, 64), (7, 130)):
        a = np.random.default_rng(1).standard_normal(shape).astype(np.float32)
        np.save(tmp_path / "in.npy", a)
        out = subprocess.run([str(exe), str(tmp_path / "in.npy"), str(tmp_path / "out.npy")], check=True, capture_output=True, text=True)
        assert out.stdout.split() == [str(shape[0]), str(shape[1])]
        b = np.load(tmp_path / "out.npy")
        assert b.dtype == np.float32 and b.shape == shape and np.array_equal(b, np.float32(2.0) * a + np.float32(1.0))


def test_cli_rejects_bad_arguments(f2d):
    app = build_app(f2d)
    assert subprocess.run([app, "--bogus"], capture_output=True).returncode == 64
    assert subprocess.run([app, "--size"], capture_output=True).returncode == 64


def test_without_gpu_fails_loudly(f2d):
    if f2d.device_count() > 0:
        pytest.skip("a GPU is present")
    r = subprocess.run([build_app(f2d), "--size", "32", "--steps", "1"], capture_output=True, text=True)
    assert r.returncode == 2 and "no CPU fallback" in r.stderr


def scripted_sources(n, s):
    """The scripted mouse of app/fluid2d_headless.cpp, restated (sources scaled by width*height in fp32)."""
    sd, su, sv = (np.zeros((n, n), np.float32) for _ in range(3))
    ph = 2.0 * math.pi * s / 97.0  # two_pi literal in the app == 2*pi in double
    ph = 6.283185307179586 * s / 97.0
    ci, cj = int(n * (0.5 + 0.25 * math.sin(ph))), int(n * (0.5 + 0.25 * math.cos(ph)))
    sd[ci, cj] += np.float32(0.075) * np.float32(n) * np.float32(n)
    vi, vj = int(n * (0.5 + 0.3 * math.sin(-1.7 * ph))), int(n * (0.5 + 0.3 * math.cos(-1.7 * ph)))
    su[vi, vj] += np.float32(0.05) * np.float32(-math.sin(-1.7 * ph)) * np.float32(n) * np.float32(n)
    sv[vi, vj] += np.float32(0.05) * np.float32(math.cos(-1.7 * ph)) * np.float32(n) * np.float32(n)
    return sd, su, sv


@pytest.mark.gpu
def test_headless_run_matches_oracle(f2d, sfo, gpu_ok, tmp_path):
    """8 frames of the app loop on a 96^2 grid (reference iteration counts 15/20, smooth on) dumped as .npy
    and compared with the oracle fed the same per-frame sources: u, v bit-exact, density within tolerance."""
    app = build_app(f2d)
    n, steps = 96, 8
    prefix = str(tmp_path / "run")
    r = subprocess.run([app, "--size", str(n), "--steps", str(steps), "--exact-divide", "--dump", prefix], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    d, u, v = (np.zeros((n, n), np.float32) for _ in range(3))
    for s in range(steps):
        sd, su, sv = scripted_sources(n, s)
        d, u, v = sfo.steps(d, sd, 0.5, u, v, su, sv, 1e-6, 0.02, 15, 20, smooth=True, sem=sfo.SEM_GPU, nsteps=1)
    gd, gu, gv = (np.load(prefix + "_%s.npy" % k) for k in ("density", "u", "v"))
    assert_bitwise(gu, u, "u")
    assert_bitwise(gv, v, "v")
    assert_close(gd, d, "density", rel_l2=2e-6 * steps, max_abs_rel=2e-5 * steps)
    # --load continues from a dumped state
    r2 = subprocess.run([app, "--size", str(n), "--steps", "0", "--load", prefix, "--dump", prefix + "_b"], capture_output=True, text=True)
    assert r2.returncode == 0, r2.stderr
    assert np.array_equal(np.load(prefix + "_b_u.npy"), gu)


@pytest.mark.gpu
def test_headless_run_cpu_exact_solver_is_bit_identical_to_fluid_solver_cpu(f2d, sfo, gpu_ok, tmp_path):
    """--solver b200-cpu-exact: the same 8 frames, now with fluid_solver_cpu's arithmetic (src/simulation.cpp:19 would
    pick fluid_solver_cpu here).  Every field bit-identical to the unmodified fluid_solver_cpu::solve when
    oracle/_ref/libref_cpu.so travelled to the box, else to the oracle pinned against it."""
    from oracle import refs

    app = build_app(f2d)
    n, steps = 96, 8
    prefix = str(tmp_path / "cpuexact")
    r = subprocess.run([app, "--size", str(n), "--steps", str(steps), "--solver", "b200-cpu-exact", "--dump", prefix],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    ref = refs.ref_cpu() if refs.have_cpu() else None
    d, u, v = (np.zeros((n, n), np.float32) for _ in range(3))
    for s in range(steps):
        sd, su, sv = scripted_sources(n, s)
        if ref is not None:
            d, u, v = ref.solve(d, sd, 0.5, u, v, su, sv, 1e-6, 0.02, 1)
        else:
            d, u, v = sfo.steps(d, sd, 0.5, u, v, su, sv, 1e-6, 0.02, 20, 20, smooth=False, sem=sfo.SEM_CPU, nsteps=1)
    for name, want in (("density", d), ("u", u), ("v", v)):
        assert_bitwise(np.load(prefix + "_%s.npy" % name), want, name)


def test_coordinates_to_cell_matches_reference_grid_renderer(f2d, tmp_path):
    """The mouse -> cell mapping (include/simulation_headless.hpp) against the UNMODIFIED grid_renderer::
    coordinates_to_cell (src/grid_renderer.cpp:3-14, compiled against oracle/sfml_stub into oracle/_ref/libref_gridr.so)."""
    import ctypes as C

    ref_path = os.path.join(ROOT, "oracle", "_ref", "libref_gridr.so")
    if not os.path.exists(ref_path):
        pytest.skip("oracle/_ref/libref_gridr.so not built (needs /root/reference)")
    src = tmp_path / "c2c.cpp"
    src.write_text('''#include "simulation_headless.hpp"
extern "C" int ours_coordinates_to_cell(size_t rows, size_t cols, float x, float y, unsigned tw, unsigned th, size_t* i, size_t* j) {
    return coordinates_to_cell(rows, cols, x, y, tw, th, *i, *j) ? 1 : 0;
}''')
    lib = tmp_path / "libc2c.so"
    libdir = os.path.join(ROOT, "fluid-2d_b200")
    subprocess.run(["g++", "-std=c++14", "-O2", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"), str(src), "-L", libdir,
                    "-lf2d", "-Wl,-rpath," + libdir, "-o", str(lib)], check=True)
    ours, ref = C.CDLL(str(lib)).ours_coordinates_to_cell, C.CDLL(ref_path).ref_coordinates_to_cell
    for fn in (ours, ref):
        fn.argtypes = [C.c_size_t, C.c_size_t, C.c_float, C.c_float, C.c_uint, C.c_uint, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
        fn.restype = C.c_int
    r = np.random.default_rng(5)
    for rows, cols, tw, th in ((800, 800, 800, 800), (256, 256, 800, 800), (100, 37, 641, 479), (4096, 4096, 1000, 3)):
        xs = np.concatenate([r.uniform(-5, tw + 5, 400), [0.0, tw - 1e-3, tw, -0.0, 0.5]]).astype(np.float32)
        ys = np.concatenate([r.uniform(-5, th + 5, 400), [0.0, th - 1e-3, th, 1.0, th / 2]]).astype(np.float32)
        for x, y in zip(xs, ys):
            a, b = (C.c_size_t(12345), C.c_size_t(54321)), (C.c_size_t(12345), C.c_size_t(54321))
            ra = ours(rows, cols, float(x), float(y), tw, th, C.byref(a[0]), C.byref(a[1]))
            rb = ref(rows, cols, float(x), float(y), tw, th, C.byref(b[0]), C.byref(b[1]))
            assert (ra, a[0].value, a[1].value) == (rb, b[0].value, b[1].value), (rows, cols, tw, th, x, y)
