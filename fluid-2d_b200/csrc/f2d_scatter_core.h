// f2d_scatter_core.h -- per-cell logic of the ordered density scatter (F2D_SEM_CPU), shared by the CUDA kernels
// (f2d_gs.cu: k_scatter_keys, k_scatter_ordered) and by the CPU harness tests/scatter_emul.cpp, like f2d_gs_tile.h.
//
// fluid_solver_cpu::advect(trace = true) (src/fluid_solver_cpu.cpp:127-152) zeroes the field, walks the sources
// (i,j) in lexicographic order and adds four weighted copies of each into the cells around its forward-traced
// position.  A target cell therefore receives its contributions in lexicographic SOURCE order, and float addition
// makes that order part of the result.  Turned inside out: every source's landing cell (i0, j0) is stored as the
// linear index key = i0 * pitch + j0 (kNoKey if the source is skipped, cpp:134); a target cell T = ti * pitch + tj
// then visits, in lexicographic order, the sources whose displacement can reach it, and a source lands on it iff
// T - key is 0, 1, pitch or pitch + 1 (j0 <= cols - 2 < pitch and tj >= 1 make the four cases unambiguous).  Only on a
// hit is the position recomputed, exactly as the reference does (product rounded, then added), and the share added.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define F2D_SC_HD __host__ __device__ __forceinline__
#else
#define F2D_SC_HD inline
#endif

namespace f2d {
namespace sc {

constexpr unsigned kNoKey = 0xffffffffu;

#if defined(__CUDA_ARCH__)
F2D_SC_HD float fadd(float a, float b) { return __fadd_rn(a, b); }
F2D_SC_HD float fsub(float a, float b) { return __fsub_rn(a, b); }
F2D_SC_HD float fmul(float a, float b) { return __fmul_rn(a, b); }
#else
F2D_SC_HD float fadd(float a, float b) { return a + b; }
F2D_SC_HD float fsub(float a, float b) { return a - b; }
F2D_SC_HD float fmul(float a, float b) { return a * b; }
#endif

// x = j + dt0*u, y = i + dt0*v (cpp:131-132); false if the source is skipped (cpp:134)
F2D_SC_HD bool forward_trace(int rows, int cols, int i, int j, float uu, float vv, float dt0, float& x, float& y) {
    x = fadd((float)j, fmul(dt0, uu));
    y = fadd((float)i, fmul(dt0, vv));
    return !(x < 0.5f || x > (float)cols - 1.5f || y < 0.5f || y > (float)rows - 1.5f);
}

// landing cell of an interior source as a linear index
F2D_SC_HD unsigned source_key(int rows, int cols, int pitch, int i, int j, float uu, float vv, float dt0) {
    float x, y;
    if (!forward_trace(rows, cols, i, j, uu, vv, dt0, x, y)) return kNoKey;
    return (unsigned)(int)y * (unsigned)pitch + (unsigned)(int)x;
}

// d = T - key; true iff the source's 2x2 footprint covers the target (kNoKey gives T + 1 >= pitch + 2: never)
F2D_SC_HD bool is_hit(unsigned d, unsigned pitch) { return !(d > pitch + 1u || (d > 1u && d < pitch)); }

// the target's share of source (i, j): (s1|s0) * (s3|s2) * value with the weights of cpp:141-149
F2D_SC_HD float share(int rows, int cols, int i, int j, float uu, float vv, float dt0, unsigned d, unsigned pitch, float value) {
    float x, y;
    forward_trace(rows, cols, i, j, uu, vv, dt0, x, y);
    const int j0 = (int)x, i0 = (int)y;  // x, y >= 0.5: truncation == the reference's static_cast<size_t>
    const float s0 = fsub(x, (float)j0), s1 = fsub(1.0f, s0), s2 = fsub(y, (float)i0), s3 = fsub(1.0f, s2);
    const float wx = (d == 1u || d == pitch + 1u) ? s0 : s1;  // landed one column left of the target: right-hand weight
    const float wy = (d >= pitch) ? s2 : s3;                  // landed one row above the target: lower weight
    return fmul(fmul(wx, wy), value);
}

// scan reach in cells from the largest displacement (as float bits); `far` = max(rows, cols) also catches NaN / inf
F2D_SC_HD int reach(unsigned disp_bits, int far) {
    union {
        unsigned u;
        float f;
    } cv;
    cv.u = disp_bits;
    const float md = cv.f;
    if (!(md < (float)far)) return far;
    int r = (int)md;
    if ((float)r < md) ++r;  // ceil
    return r + 1;
}

}  // namespace sc
}  // namespace f2d
