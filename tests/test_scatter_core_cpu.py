"""CPU check of the ordered density scatter's per-cell logic (fluid-2d_b200/csrc/f2d_scatter_core.h, shared with the
CUDA kernels k_scatter_keys / k_scatter_ordered): landing-cell keys, the hit rule and the weights, executed on the host
by tests/scatter_emul.cpp, against the oracle's sequential scatter of fluid_solver_cpu
(src/fluid_solver_cpu.cpp:127-152).  Interior cells bit for bit (edges and corners belong to the boundary pass)."""
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
    src = os.path.join(HERE, "scatter_emul.cpp")
    hdr = os.path.join(ROOT, "fluid-2d_b200", "csrc", "f2d_scatter_core.h")
    out_dir = os.path.join(HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    lib = os.path.join(out_dir, "libscatter_emul.so")
    if not os.path.exists(lib) or os.path.getmtime(lib) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-o", lib, src],
                       check=True)
    L = C.CDLL(lib)
    L.scatter_emul.argtypes = [_FP, _FP, _FP, _FP, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int]
    L.scatter_emul.restype = None
    return L


@pytest.mark.parametrize("rows,cols,cells,seed", [(3, 3, 1.0, 1), (16, 16, 2.0, 2), (40, 70, 3.5, 3), (64, 64, 0.3, 4),
                                                   (50, 33, 9.0, 5), (48, 48, 60.0, 6)])
def test_keyed_gather_equals_sequential_scatter(sfo, emul, rows, cols, cells, seed):
    d, u, v, *_ = rng_fields(rows, 3000 + seed, cols=cols, vel_cells=cells)
    want = sfo.advect_scatter(d, u, v, sfo.BND_CONTINUOUS, DT, sfo.SEM_CPU)
    dt0 = np.float32(np.sqrt(np.float32(rows * cols))) * np.float32(DT)  # cpp:126
    pitch = (cols + 31) // 32 * 32

    def pad(a):
        b = np.zeros((rows, pitch), np.float32)
        b[:, :cols] = a
        return b

    for rev in (0, 1):
        out = np.full((rows, pitch), np.float32(-7.0))
        emul.scatter_emul(pad(d).ctypes.data_as(_FP), pad(u).ctypes.data_as(_FP), pad(v).ctypes.data_as(_FP),
                          out.ctypes.data_as(_FP), rows, cols, pitch, float(dt0), rev)
        assert_bitwise(np.ascontiguousarray(out[1:-1, 1:cols - 1]), np.ascontiguousarray(want[1:-1, 1:-1]),
                       "interior %dx%d, displacement ~%.1f cells, order %d" % (rows, cols, cells, rev))
        assert np.all(out[:, cols:] == np.float32(-7.0)), "padding columns were written"
