"""CPU check of the Gauss-Seidel wavefront's tile core (fluid-2d_b200/csrc/f2d_gs_tile.h).

The CUDA kernel k_gs_relax calls the header's per-lane functions; tests/gs_emul.cpp compiles the same
header with g++ and executes the tiles of all sweeps in random orders constrained only by the kernel's own
wait conditions.  Every order must reproduce the oracle's sequential in-place sweeps of
fluid_solver_cpu (src/fluid_solver_cpu.cpp:104-113, :196-204) bit for bit, on every cell but the four
corners (which the product averages in a separate kernel after the last sweep).
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from util import DT, assert_bitwise, rng_fields

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
_FP = C.POINTER(C.c_float)


@pytest.fixture(scope="module")
def emul():
    src = os.path.join(HERE, "gs_emul.cpp")
    hdr = os.path.join(ROOT, "fluid-2d_b200", "csrc", "f2d_gs_tile.h")
    out_dir = os.path.join(HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    lib = os.path.join(out_dir, "libgs_emul.so")
    if not os.path.exists(lib) or os.path.getmtime(lib) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
                        "-o", lib, src], check=True)
    L = C.CDLL(lib)
    L.gs_emul_relax.argtypes = [_FP, _FP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                                C.c_int, C.c_uint64]
    L.gs_emul_relax.restype = C.c_long
    return L


def relax(L, x, rhs, kind, diffuse, a, c, K, order, pitch=None):
    rows, cols = x.shape
    pitch = pitch or cols
    xb = np.full((rows, pitch), np.float32(777.0))
    rb = np.full((rows, pitch), np.float32(-777.0))
    xb[:, :cols] = x
    rb[:, :cols] = rhs
    blocked = L.gs_emul_relax(xb.ctypes.data_as(_FP), rb.ctypes.data_as(_FP), rows, cols, pitch, kind, int(diffuse),
                              a, c, K, order)
    assert blocked >= 0, "the sequential schedule blocked: wait conditions are not implied by program order"
    assert np.all(xb[:, cols:] == np.float32(777.0)), "padding columns were written"
    return np.ascontiguousarray(xb[:, :cols]), blocked


def same_but_corners(a, ref, what):
    a, ref = a.copy(), ref.copy()
    for i in (0, -1):
        for j in (0, -1):
            a[i, j] = ref[i, j] = 0.0
    assert_bitwise(a, ref, what)


@pytest.mark.parametrize("rows,cols,K,seed", [(3, 3, 2, 1), (5, 70, 3, 2), (34, 34, 4, 3), (35, 67, 5, 4),
                                              (70, 40, 6, 5), (100, 100, 8, 6), (130, 96, 20, 7)])
@pytest.mark.parametrize("rate", [0.5, 1e-6])
def test_diffuse_tiles_any_order(sfo, emul, rows, cols, K, seed, rate):
    f = rng_fields(rows, seed, cols=cols)[1]  # a signed field
    a = np.float32(np.float32(np.float32(DT) * np.float32(rows * cols)) * np.float32(rate))
    c = np.float32(np.float32(1.0) + np.float32(4.0) * a)
    for kind in (sfo.BND_CONTINUOUS, sfo.BND_OPPOSITE_HORIZONTAL, sfo.BND_OPPOSITE_VERTICAL):
        want = sfo.diffuse(f, kind, rate, DT, K, sem=sfo.SEM_CPU)
        pitch = (cols + 31) // 32 * 32
        for order in (0, 11 + seed, 1234567 + kind):
            got, blocked = relax(emul, f, f, kind, True, a, c, K, order, pitch=pitch)
            same_but_corners(got, want, "diffuse %dx%d K=%d kind=%d order=%d (blocked %d)" % (rows, cols, K, kind, order, blocked))


@pytest.mark.parametrize("n,K,seed", [(16, 3, 1), (33, 7, 2), (64, 20, 3), (97, 12, 4)])
def test_pressure_tiles_any_order(sfo, emul, n, K, seed):
    _, u, v, _, _, _ = rng_fields(n, seed)
    _, _, p_want, dv = sfo.project(u, v, K, sem=sfo.SEM_CPU, return_p=True)
    for order in (0, 99 + seed):
        got, _ = relax(emul, np.zeros_like(dv), dv, sfo.BND_CONTINUOUS, False, 0.0, 1.0, K, order)
        same_but_corners(got, p_want, "pressure %d K=%d order=%d" % (n, K, order))


def test_random_orders_are_adversarial(sfo, emul):
    """The random scheduler really does run sweeps out of order: it must find bands blocked."""
    f = rng_fields(70, 9)[1]
    a = np.float32(np.float32(np.float32(DT) * np.float32(70 * 70)) * np.float32(0.5))
    _, blocked = relax(emul, f, f, 0, True, a, np.float32(1.0) + np.float32(4.0) * a, 6, 4242)
    assert blocked > 0
