"""Canonical synthetic benchmark inputs: ctypes front end of tools/canonical_fields.c.

Neutral workload generator -- not product code, not the oracle: bench.py's product arm, bench_multi.py and the
tools use it so that they never load anything under oracle/.  tests/test_canonical_fields.py pins it bit for bit to
the oracle's own generator (and through it to the FNV anchors of SURVEY.md Appendix D)."""
import ctypes as C
import os
import subprocess
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "canonical_fields.c")
_LIB = os.path.join(_HERE, "libcanonical.so")
_lib = None
_FP = C.POINTER(C.c_float)


def build(force=False):
    if force or not os.path.exists(_LIB) or (os.path.exists(_SRC) and os.path.getmtime(_LIB) < os.path.getmtime(_SRC)):
        subprocess.run(["gcc", "-std=c11", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", _LIB, _SRC, "-lm"], check=True)
    return _LIB


def _load():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        L.f2d_canonical_fields.argtypes = [C.c_size_t, C.c_size_t, C.c_size_t] + [_FP] * 6
        L.f2d_canonical_fields.restype = None
        _lib = L
    return _lib


def rows(n, r0, r1, threads=None):
    """Rows [r0, r1) of the six canonical N x N fields (d, u, v, sd, su, sv) as (r1-r0, n) float32 arrays,
    generated locally: a slab never needs the full-grid host arrays."""
    L = _load()
    nrows = r1 - r0
    arrs = [np.empty((nrows, n), dtype=np.float32) for _ in range(6)]
    # the generator addresses cells at their GLOBAL position: shift every base pointer back by r0 rows
    ptrs = [C.cast(a.ctypes.data - r0 * n * 4, _FP) for a in arrs]
    threads = threads or min(32, os.cpu_count() or 1, max(1, nrows // 256))
    bounds = np.linspace(r0, r1, threads + 1).astype(int)
    ts = [threading.Thread(target=L.f2d_canonical_fields, args=(n, int(bounds[k]), int(bounds[k + 1]), *ptrs))
          for k in range(threads)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    return tuple(arrs)


def fields(n, threads=None):
    """(d, u, v, sd, su, sv) on the full N x N grid."""
    return rows(n, 0, n, threads)
