"""ctypes binding of include/f2d.h (libf2d.so).  Host-side plumbing only: every compute call
goes through the C ABI into the CUDA kernels; nothing here computes on the CPU."""
import ctypes as C
import os
import re
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_LIB = os.environ.get("F2D_LIB_PATH") or os.path.join(_HERE, "libf2d.so")  # F2D_LIB_PATH: A/B kernel variants
_HEADER = os.path.join(_ROOT, "include", "f2d.h")

OK, ERR_INVALID, ERR_CUDA, ERR_NO_DEVICE, ERR_STATE = 0, 1, 2, 3, 4
FIELD_DENSITY, FIELD_U, FIELD_V = 0, 1, 2
FIELD_DENSITY_SOURCE, FIELD_U_SOURCE, FIELD_V_SOURCE = 3, 4, 5
FIELD_PRESSURE, FIELD_DIVERGENCE = 6, 7
BND_CONTINUOUS, BND_OPPOSITE_HORIZONTAL, BND_OPPOSITE_VERTICAL = 0, 1, 2
JACOBI_NAIVE, JACOBI_STREAM = 0, 1
DIV_F64, DIV_F32_CORR = 0, 1
SEM_GPU, SEM_CPU = 0, 1


class F2DError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("f2d error %d: %s" % (code, msg))
        self.code = code


class SolverConfig(C.Structure):
    """struct f2d_config (include/f2d.h)."""

    _fields_ = [
        ("struct_size", C.c_uint32),
        ("rows", C.c_uint32),
        ("cols", C.c_uint32),
        ("diffuse_iters", C.c_uint32),
        ("project_iters", C.c_uint32),
        ("smooth", C.c_uint32),
        ("jacobi_mode", C.c_uint32),
        ("temporal_block", C.c_uint32),
        ("divide_mode", C.c_uint32),
        ("use_graph", C.c_uint32),
        ("device", C.c_int32),
        ("global_rows", C.c_uint32),
        ("row_offset", C.c_uint32),
        ("halo", C.c_uint32),
        ("temporal_block_diffuse", C.c_uint32),
        ("semantics", C.c_uint32),
        ("stream", C.c_void_p),
    ]


def lib_path():
    return _LIB


def build(force=False):
    """Compile libf2d.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    srcdir = os.path.join(_HERE, "csrc")
    if not force and os.path.exists(_LIB):
        newest = max(os.path.getmtime(os.path.join(srcdir, f)) for f in os.listdir(srcdir)
                     if f.endswith((".cu", ".cuh", ".h")) or f == "Makefile")
        newest = max(newest, os.path.getmtime(_HEADER))
        if os.path.getmtime(_LIB) >= newest:
            return _LIB
    subprocess.run(["make", "-C", srcdir, "-j8", "all"], check=True, stdout=subprocess.DEVNULL)
    return _LIB


def abi_symbols():
    """Names of every function include/f2d.h declares (for the export check in the CPU tests)."""
    with open(_HEADER) as f:
        text = f.read()
    return sorted(set(re.findall(r"F2D_API\s+[\w\s\*]+?\b(f2d_\w+)\s*\(", text)))


_lib = None


def load():
    """dlopen libf2d.so; raises (never falls back) when the CUDA library is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB):
        raise F2DError(ERR_STATE, "libf2d.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`);"
                                  " there is no CPU fallback")
    L = C.CDLL(_LIB)
    u32, f32, i32, vp = C.c_uint32, C.c_float, C.c_int, C.c_void_p
    fp = C.POINTER(C.c_float)
    cfgp = C.POINTER(SolverConfig)
    sig = {
        "f2d_config_default": [cfgp, u32, u32],
        "f2d_create": [cfgp, C.POINTER(vp)],
        "f2d_solve_host": [vp, fp, fp, f32, fp, fp, fp, fp, f32, f32],
        "f2d_upload": [vp, fp, fp, fp],
        "f2d_set_sources": [vp, fp, fp, fp],
        "f2d_download": [vp, fp, fp, fp],
        "f2d_upload_field": [vp, i32, fp],
        "f2d_download_field": [vp, i32, fp],
        "f2d_clear_sources": [vp],
        "f2d_step": [vp, f32, f32, f32, u32],
        "f2d_step_timed": [vp, f32, f32, f32, u32, C.POINTER(f32)],
        "f2d_sync": [vp],
        "f2d_stage_set_bnd": [vp, i32, i32],
        "f2d_stage_add_sources": [vp, i32, f32],
        "f2d_stage_diffuse": [vp, i32, i32, f32, f32, u32],
        "f2d_stage_smooth": [vp],
        "f2d_stage_advect_density": [vp, f32],
        "f2d_stage_advect_velocity": [vp, f32],
        "f2d_stage_project": [vp, u32],
        "f2d_bench_jacobi": [vp, i32, u32, u32, C.POINTER(f32)],
        "f2d_launch_count": [vp, C.POINTER(C.c_uint64)],
        "f2d_field_ptr": [vp, i32, C.POINTER(vp), C.POINTER(C.c_size_t)],
        "f2d_get_config": [vp, cfgp],
        "f2d_get_stream": [vp, C.POINTER(vp)],
        "f2d_comm_unique_id": [C.c_char_p],
        "f2d_comm_init": [vp, C.c_char_p, i32, i32, i32],
        "f2d_comm_stats": [vp, C.POINTER(C.c_uint64)],
        "f2d_comm_bytes": [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)],
        "f2d_pin_host": [vp, vp, C.c_size_t],
        "f2d_unpin_host": [vp, vp],
        "f2d_p2p_export": [vp, C.POINTER(C.c_ubyte), C.POINTER(C.c_uint64)],
        "f2d_p2p_connect": [vp, i32, i32, C.POINTER(C.c_ubyte), C.POINTER(C.c_uint64), C.POINTER(C.c_ubyte),
                            C.POINTER(C.c_uint64), i32],
        "f2d_render_density_rgba": [vp, f32, f32, f32, C.POINTER(C.c_ubyte)],
        "f2d_render_velocity_lines": [vp, f32, f32, fp],
        "f2d_abi_version": [],
        "f2d_device_count": [],
    }
    for name, args in sig.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = C.c_int
    L.f2d_destroy.argtypes = [vp]
    L.f2d_destroy.restype = None
    L.f2d_last_error.argtypes = []
    L.f2d_last_error.restype = C.c_char_p
    _lib = L
    return L


def check(rc):
    if rc != OK:
        raise F2DError(rc, load().f2d_last_error().decode("utf-8", "replace"))


def device_count():
    return int(load().f2d_device_count())
