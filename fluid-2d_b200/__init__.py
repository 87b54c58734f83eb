"""fluid2d_b200 -- B200-native stable-fluids step behind the fluid-2d solver interface.

The product is the CUDA library `libf2d.so` (C ABI: include/f2d.h, sources: csrc/).  This
package is the thin Python host mirror used by the tests and bench.py:

    FluidSolverB200   mirror of the reference's `fluid_solver` interface
                      (src/fluid_solver.hpp:8-25): solve(density, density_source, diffusion_rate,
                      u, v, u_source, v_source, viscosity, dt) on host arrays, in place.
    SolverConfig      f2d_config.
    slab              row-slab decomposition helpers for the multi-GPU path.

There is NO CPU fallback: importing works anywhere, but creating a solver without the built
library or without a CUDA device raises.
"""
from .capi import (  # noqa: F401
    BND_CONTINUOUS,
    BND_OPPOSITE_HORIZONTAL,
    BND_OPPOSITE_VERTICAL,
    DIV_F32_CORR,
    DIV_F64,
    FIELD_DENSITY,
    FIELD_DENSITY_SOURCE,
    FIELD_DIVERGENCE,
    FIELD_PRESSURE,
    FIELD_U,
    FIELD_U_SOURCE,
    FIELD_V,
    FIELD_V_SOURCE,
    JACOBI_NAIVE,
    JACOBI_STREAM,
    SEM_CPU,
    SEM_GPU,
    F2DError,
    SolverConfig,
    abi_symbols,
    build,
    device_count,
    lib_path,
    load,
)
from . import slab  # noqa: F401
from .solver import FluidSolverB200  # noqa: F401

__all__ = ["FluidSolverB200", "SolverConfig", "F2DError", "build", "load", "device_count"]
