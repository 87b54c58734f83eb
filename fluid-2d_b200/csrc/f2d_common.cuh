// f2d_common.cuh -- shared device/host definitions for the stable-fluids kernels (sm_100a).
//
// Replaces the reference's buffer layer (element_accessor / gpu_buffer / linear_buffer,
// src/gpu_buffer.hpp:9-63, src/linear_buffer.hpp:11-49): a device field is a plain fp32 array
// whose rows start 128-byte aligned (pitch is a multiple of 32 floats), addressed with 32-bit
// index math, so that whole rows can be moved with 16-byte vector loads / async copies.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/f2d.h"

namespace f2d {

// Geometry of the LOCAL field held by one solver (one row slab of the global grid).
struct Geom {
    int rows;   // local rows (incl. halo rows of a slab)
    int cols;   // columns (global == local: slabs split rows only)
    int pitch;  // floats between consecutive rows (multiple of 32)
    int grow0;  // global row index of local row 0
    int grows;  // global rows
};

// Coefficients of one diffuse solve: a = dt*float(rows*cols)*rate (src/fluid_solver_gpu.cu:79).
struct DiffuseCoef {
    float a;
    float rc, rl;  // 1/c split as rc + rl (rc = RN32(1/c), rl = RN32(1/c - rc)): 48 bits of the reciprocal
    double c;      // 1.0 + 4.0*(double)a, the reference's fp64 divisor (gpu.cu:82)
};

// ---- cell classification for the fused boundary pass ------------------------------------
// The reference runs set_boundary_* as separate 1-D kernels after every 2-D kernel
// (src/fluid_solver_gpu.cu:260-276).  Every edge value is +/- the adjacent interior value of the
// same iterate and corners are never written (gpu.cu:15-23), so each kernel here produces the
// edge cells itself: an edge thread evaluates the update of its inward neighbour and applies
// the sign; corner threads (and the first/last local row of a slab that is not a global edge)
// pass the input value through.
enum CellClass { CELL_COMPUTE = 0, CELL_KEEP = 1 };

struct CellSrc {
    int cls;  // CELL_COMPUTE: value = sign * update(si, sj);  CELL_KEEP: copy input(i, j)
    int si, sj;
    bool negate;
};

__device__ __forceinline__ CellSrc classify_cell(const Geom& g, int i, int j, int kind) {
    CellSrc c;
    c.cls = CELL_COMPUTE;
    c.si = i;
    c.sj = j;
    c.negate = false;
    const int gi = g.grow0 + i;
    const bool top = (gi == 0), bottom = (gi == g.grows - 1);
    const bool left = (j == 0), right = (j == g.cols - 1);
    if ((top || bottom) && (left || right)) {  // corner: never written by the reference
        c.cls = CELL_KEEP;
        return c;
    }
    if ((i == 0 && !top) || (i == g.rows - 1 && !bottom)) {  // slab-local edge row: halo, not computed
        c.cls = CELL_KEEP;
        return c;
    }
    if (left) {
        c.sj = 1;
        c.negate = (kind == F2D_BND_OPPOSITE_HORIZONTAL);
    } else if (right) {
        c.sj = g.cols - 2;
        c.negate = (kind == F2D_BND_OPPOSITE_HORIZONTAL);
    } else if (top) {
        c.si = i + 1;
        c.negate = (kind == F2D_BND_OPPOSITE_VERTICAL);
    } else if (bottom) {
        c.si = i - 1;
        c.negate = (kind == F2D_BND_OPPOSITE_VERTICAL);
    }
    // a mirrored source on a slab-local edge row cannot be evaluated (no neighbour rows)
    if ((c.si == 0 && g.grow0 != 0) || (c.si == g.rows - 1 && g.grow0 + g.rows != g.grows)) c.cls = CELL_KEEP;
    return c;
}

// ---- arithmetic, pinned with intrinsics (the library is compiled with -fmad=false) ---------
// The reference is compiled with nvcc defaults, i.e. FMA contraction wherever ptxas found it;
// the SASS of each reference kernel was read and the same operations are
// spelled out here so that results do not depend on this compiler's contraction choices.

// diffuse_iteration_kernel (src/fluid_solver_gpu.cu:81-82):
//   num = FFMA(a, ((W + E) + N) + S, x0);  out = (float)((double)num / (1.0 + 4.0*a))
template <int DIVMODE>
__device__ __forceinline__ float diffuse_update(float w, float e, float n, float s, float x0,
                                                const DiffuseCoef& k) {
    float sum = __fadd_rn(__fadd_rn(__fadd_rn(w, e), n), s);
    float num = __fmaf_rn(k.a, sum, x0);
    if (DIVMODE == F2D_DIV_F64) {
        return __double2float_rn(__ddiv_rn((double)num, k.c));
    } else {
        // Division by a constant as a correctly rounded multiplication by the 48-bit constant 1/c = rc + rl
        // (Brisebarre & Muller, "Correctly rounded multiplication by arbitrary precision constants"): the low product
        // is rounded once, the high product and the sum are exact inside the FMA, so the result is RN32(num / c)
        // unless num / c lies within ~2^-24 ulp of a rounding midpoint (measured: 1 of 2.2e8 random quotients over the
        // coefficients of all published grids differs from the fp64 divide, by one ulp).  Two operations instead of
        // the fp64 divide's ~40 and of the four of round 1's residual correction (same accuracy).
        return __fmaf_rn(num, k.rc, __fmul_rn(num, k.rl));
    }
}

// p_iteration_kernel (src/fluid_solver_gpu.cu:187-188): ((((div + pE) + pW) + pS) + pN) * 0.25f
__device__ __forceinline__ float pressure_update(float dv, float e, float w, float s, float n) {
    return __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(dv, e), w), s), n), 0.25f);
}

// calculate_divergence_kernel (gpu.cu:173-174): (-0.5f*h) * (((uE - uW) + vS) - vN)
__device__ __forceinline__ float divergence_update(float ue, float uw, float vs, float vn, float mhalf_h) {
    return __fmul_rn(mhalf_h, __fsub_rn(__fadd_rn(__fsub_rn(ue, uw), vs), vn));
}

// remove_p_kernel (gpu.cu:202-203): f - (0.5f*(p_hi - p_lo)) / h   (IEEE divide, then subtract)
__device__ __forceinline__ float gradient_update(float f, float p_hi, float p_lo, float h) {
    return __fsub_rn(f, __fdiv_rn(__fmul_rn(0.5f, __fsub_rn(p_hi, p_lo)), h));
}

// smooth_kernel (gpu.cu:94): 0.2f * ((((c + W) + E) + N) + S)
__device__ __forceinline__ float smooth_update(float c, float w, float e, float n, float s) {
    return __fmul_rn(0.2f, __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(c, w), e), n), s));
}

// bilinear weights of advect_kernel / advect_trace_kernel (gpu.cu:116-123, :146-154)
struct Bilinear {
    int i0, j0;
    float s0, s1, s2, s3;
};
__device__ __forceinline__ Bilinear bilinear_setup(float x, float y) {
    Bilinear b;
    b.j0 = (int)x;  // x,y >= 0.5 here: truncation == the reference's static_cast<size_t>
    b.i0 = (int)y;
    b.s0 = __fsub_rn(x, (float)b.j0);
    b.s1 = __fsub_rn(1.0f, b.s0);
    b.s2 = __fsub_rn(y, (float)b.i0);
    b.s3 = __fsub_rn(1.0f, b.s2);
    return b;
}
// advect_kernel (gpu.cu:125-126) as contracted by nvcc: FMUL,FMUL,FFMA,FMUL,FFMA,FFMA
__device__ __forceinline__ float bilinear_gather(const Bilinear& b, float a00, float a01, float a10, float a11) {
    float top = __fmaf_rn(b.s0, a01, __fmul_rn(b.s1, a00));
    float bot = __fmaf_rn(b.s0, a11, __fmul_rn(b.s1, a10));
    return __fmaf_rn(b.s3, top, __fmul_rn(b.s2, bot));
}

__device__ __forceinline__ float apply_sign(float v, bool negate) { return negate ? -v : v; }

}  // namespace f2d
