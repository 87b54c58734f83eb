"""CPU-side checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, exports
every symbol include/f2d.h declares, and FAILS LOUDLY without a GPU (no CPU fallback)."""
import ctypes as C
import os
import subprocess

import pytest


def test_library_builds_and_exports_every_declared_symbol(f2d):
    path = f2d.lib_path()
    assert os.path.exists(path)
    declared = f2d.abi_symbols()
    assert len(declared) >= 25
    out = subprocess.run(["nm", "-D", "--defined-only", path], check=True, capture_output=True, text=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    missing = [s for s in declared if s not in exported]
    assert not missing, "declared in include/f2d.h but not exported: %r" % missing
    extra = sorted(s for s in exported if s.startswith("f2d_") and s not in declared)
    assert not extra, "exported but not declared: %r" % extra


def test_library_contains_sm100a_code(f2d):
    out = subprocess.run(["cuobjdump", "-lelf", f2d.lib_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_abi_version_and_default_config(f2d):
    L = f2d.load()
    assert L.f2d_abi_version() == 3
    cfg = f2d.SolverConfig()
    assert L.f2d_config_default(C.byref(cfg), 256, 256) == 0
    assert cfg.struct_size == C.sizeof(f2d.SolverConfig)
    # defaults == fluid_solver_gpu::solve literals (src/fluid_solver_gpu.cu:238-252)
    assert (cfg.diffuse_iters, cfg.project_iters, cfg.smooth) == (15, 20, 1)
    assert cfg.jacobi_mode == f2d.JACOBI_STREAM and cfg.global_rows == 256
    assert cfg.semantics == f2d.SEM_GPU  # fluid_solver_gpu arithmetic unless asked otherwise


def test_create_without_gpu_raises_no_fallback(f2d):
    if f2d.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(f2d.F2DError) as ei:
        f2d.FluidSolverB200(64, 64)
    assert ei.value.code == 3 and "no CPU fallback" in str(ei.value)


def test_bad_arguments_are_rejected(f2d):
    L = f2d.load()
    cfg = f2d.SolverConfig()
    L.f2d_config_default(C.byref(cfg), 2, 2)
    h = C.c_void_p()
    assert L.f2d_create(C.byref(cfg), C.byref(h)) == 1  # F2D_ERR_INVALID: grid too small
    cfg.struct_size = 4
    assert L.f2d_create(C.byref(cfg), C.byref(h)) == 1
    assert L.f2d_step(None, 0.0, 0.0, 0.02, 1) == 1
    assert b"NULL" in L.f2d_last_error()


def test_product_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under fluid-2d_b200/ or include/ may reference it."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for base in ("fluid-2d_b200", "include"):
        for dp, _, files in os.walk(os.path.join(root, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                    text = open(os.path.join(dp, f), errors="replace").read()
                    assert "oracle" not in text.lower() or f == "f2d_common.cuh" and False, os.path.join(dp, f)
