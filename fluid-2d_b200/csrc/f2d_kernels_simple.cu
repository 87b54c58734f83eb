// f2d_kernels_simple.cu -- the one-pass stages of the stable-fluids step for sm_100a.
//
// Each kernel is one thread per cell over the FULL local field (edges included) and produces
// its own boundary cells (see classify_cell), so no separate set_boundary launches and no
// full-field copies are needed (the reference issues 94 + 94 of them per step,
// SURVEY.md Appendix C).  All kernels are out-of-place unless noted.
#include "f2d_kernels.cuh"

namespace f2d {

namespace {
constexpr int kBx = 32, kBy = 8;
inline dim3 grid2d(const Geom& g, int z = 1) { return dim3((g.cols + kBx - 1) / kBx, (g.rows + kBy - 1) / kBy, z); }
#define F2D_CELL_IJ()                                   \
    const int j = blockIdx.x * kBx + threadIdx.x;       \
    const int i = blockIdx.y * kBy + threadIdx.y;       \
    if (i >= g.rows || j >= g.cols) return;
}  // namespace

// ---------------------------------------------------------------------------- add_sources
// add_sources_kernel (src/fluid_solver_gpu.cu:56-67): o = FFMA(dt, s, f) on the global interior, no
// boundary pass; every other cell is passed through, so the destination may be the field itself
// (in place) or a fresh buffer.  Up to three fields per launch (blockIdx.z).
__global__ void __launch_bounds__(kBx* kBy) k_add_sources(Geom g, AddSourceBatch b, float dt) {
    F2D_CELL_IJ();
    const int gi = g.grow0 + i;
    const size_t o = (size_t)i * g.pitch + j;
    const float* f = b.f[blockIdx.z];
    float* out = b.o[blockIdx.z];
    const bool interior = !(gi < 1 || gi > g.grows - 2 || j < 1 || j > g.cols - 2);
    if (interior)
        out[o] = __fmaf_rn(dt, __ldg(b.s[blockIdx.z] + o), f[o]);
    else if (out != f)
        out[o] = f[o];
}

// ---- float4 row kernels (cols % 4 == 0): one thread = four consecutive cells of one row ----------
// Columns 0/1 and cols-2/cols-1 then share a float4, so the fused boundary pass needs no second
// evaluation: the edge column takes +/- its neighbour component, an edge ROW is produced by
// evaluating the adjacent interior row, corners pass the input through.
namespace {
constexpr int kVx = 64, kVy = 4;
inline dim3 grid_v4(const Geom& g, int z = 1) { return dim3((g.cols / 4 + kVx - 1) / kVx, (g.rows + kVy - 1) / kVy, z); }

struct RowSrc {
    int si;        // row whose update is evaluated (== i for interior rows)
    bool keep;     // pass the input row through (slab-local edge row)
    bool edge_row; // i is a global top/bottom row
};
__device__ __forceinline__ RowSrc classify_row(const Geom& g, int i) {
    RowSrc r;
    const int gi = g.grow0 + i;
    const bool top = (gi == 0), bottom = (gi == g.grows - 1);
    r.edge_row = top || bottom;
    r.si = top ? i + 1 : (bottom ? i - 1 : i);
    r.keep = ((i == 0 && !top) || (i == g.rows - 1 && !bottom));
    if ((r.si == 0 && g.grow0 != 0) || (r.si == g.rows - 1 && g.grow0 + g.rows != g.grows)) r.keep = true;
    return r;
}
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 neg4(const float4& v, bool n) {
    return n ? make_float4(-v.x, -v.y, -v.z, -v.w) : v;
}
#define F2D_ROW_J4()                                        \
    const int jg = blockIdx.x * kVx + threadIdx.x;          \
    const int i = blockIdx.y * kVy + threadIdx.y;           \
    const int j = 4 * jg;                                   \
    if (i >= g.rows || j >= g.cols) return;
}  // namespace

__global__ void __launch_bounds__(kVx* kVy) k_add_sources_v4(Geom g, AddSourceBatch b, float dt) {
    F2D_ROW_J4();
    const int gi = g.grow0 + i;
    const size_t o = (size_t)i * g.pitch + j;
    const float* __restrict__ f = b.f[blockIdx.z];
    float* out = b.o[blockIdx.z];
    const float4 fv = ld4(f + o);
    if (gi < 1 || gi > g.grows - 2) {
        if (out != f) st4(out + o, fv);
        return;
    }
    const float4 sv = ld4(b.s[blockIdx.z] + o);
    float4 r;
    r.x = (j == 0) ? fv.x : __fmaf_rn(dt, sv.x, fv.x);
    r.y = __fmaf_rn(dt, sv.y, fv.y);
    r.z = __fmaf_rn(dt, sv.z, fv.z);
    r.w = (j + 3 == g.cols - 1) ? fv.w : __fmaf_rn(dt, sv.w, fv.w);
    st4(out + o, r);
}

void launch_add_sources(const Geom& g, const AddSourceBatch& b, float dt, cudaStream_t st) {
    if (g.cols % 4 == 0)
        k_add_sources_v4<<<grid_v4(g, b.n), dim3(kVx, kVy), 0, st>>>(g, b, dt);
    else
        k_add_sources<<<grid2d(g, b.n), dim3(kBx, kBy), 0, st>>>(g, b, dt);
}

// --------------------------------------------------------------------------- naive Jacobi
// One sweep of diffuse_iteration_kernel (gpu.cu:69-85) or p_iteration_kernel (gpu.cu:179-191)
// with the boundary pass fused; ping-pong buffers replace the reference's per-iteration
// full-field copy (gpu.cu:302, :380).  Bring-up / cross-check path (F2D_JACOBI_NAIVE).
template <bool DIFFUSE, int DIVMODE>
__global__ void __launch_bounds__(kBx* kBy) k_jacobi_naive(Geom g, RelaxBatch b) {
    F2D_CELL_IJ();
    const RelaxField& fld = b.f[blockIdx.z];
    const CellSrc c = classify_cell(g, i, j, fld.kind);
    const size_t o = (size_t)i * g.pitch + j;
    const float* __restrict__ prev = fld.prev;
    if (c.cls == CELL_KEEP) {
        fld.next[o] = prev ? prev[o] : 0.0f;
        return;
    }
    const size_t so = (size_t)c.si * g.pitch + c.sj;
    float W = 0.f, E = 0.f, N = 0.f, S = 0.f;
    if (prev) {
        W = prev[so - 1];
        E = prev[so + 1];
        N = prev[so - g.pitch];
        S = prev[so + g.pitch];
    }
    const float r = __ldg(fld.rhs + so);
    float val;
    if (DIFFUSE)
        val = diffuse_update<DIVMODE>(W, E, N, S, r, fld.coef);
    else
        val = pressure_update(r, E, W, S, N);
    fld.next[o] = apply_sign(val, c.negate);
}

void launch_jacobi_naive(const Geom& g, const RelaxBatch& b, bool diffuse, int divmode, cudaStream_t st) {
    dim3 gr = grid2d(g, b.n), bl(kBx, kBy);
    if (!diffuse)
        k_jacobi_naive<false, F2D_DIV_F64><<<gr, bl, 0, st>>>(g, b);
    else if (divmode == F2D_DIV_F64)
        k_jacobi_naive<true, F2D_DIV_F64><<<gr, bl, 0, st>>>(g, b);
    else
        k_jacobi_naive<true, F2D_DIV_F32_CORR><<<gr, bl, 0, st>>>(g, b);
}

// ----------------------------------------------------------------------------- divergence
// calculate_divergence_kernel (gpu.cu:164-177) + set_boundary_continuous (gpu.cu:376); corners are
// the zeros of the reference's memset (gpu.cu:363).
__global__ void __launch_bounds__(kBx* kBy) k_divergence(Geom g, const float* __restrict__ u,
                                                        const float* __restrict__ v, float* __restrict__ dv,
                                                        float mhalf_h) {
    F2D_CELL_IJ();
    const CellSrc c = classify_cell(g, i, j, F2D_BND_CONTINUOUS);
    const size_t o = (size_t)i * g.pitch + j;
    if (c.cls == CELL_KEEP) {
        dv[o] = 0.0f;
        return;
    }
    const size_t so = (size_t)c.si * g.pitch + c.sj;
    dv[o] = divergence_update(u[so + 1], u[so - 1], v[so + g.pitch], v[so - g.pitch], mhalf_h);
}

__global__ void __launch_bounds__(kVx* kVy) k_divergence_v4(Geom g, const float* __restrict__ u,
                                                           const float* __restrict__ v, float* __restrict__ dv,
                                                           float mhalf_h) {
    F2D_ROW_J4();
    const RowSrc rs = classify_row(g, i);
    const size_t o = (size_t)i * g.pitch + j;
    if (rs.keep) {
        st4(dv + o, make_float4(0.f, 0.f, 0.f, 0.f));
        return;
    }
    const size_t so = (size_t)rs.si * g.pitch + j;
    const float4 uc = ld4(u + so), vs = ld4(v + so + g.pitch), vn = ld4(v + so - g.pitch);
    const float ul = (j > 0) ? __ldg(u + so - 1) : 0.f;
    const float ur = (j + 4 < g.cols) ? __ldg(u + so + 4) : 0.f;
    float4 r;
    r.x = divergence_update(uc.y, ul, vs.x, vn.x, mhalf_h);
    r.y = divergence_update(uc.z, uc.x, vs.y, vn.y, mhalf_h);
    r.z = divergence_update(uc.w, uc.y, vs.z, vn.z, mhalf_h);
    r.w = divergence_update(ur, uc.z, vs.w, vn.w, mhalf_h);
    // set_boundary_continuous: edge columns copy their neighbour, corners stay 0 (the memset of gpu.cu:363)
    if (j == 0) r.x = rs.edge_row ? 0.f : r.y;
    if (j + 3 == g.cols - 1) r.w = rs.edge_row ? 0.f : r.z;
    st4(dv + o, r);
}

void launch_divergence(const Geom& g, const float* u, const float* v, float* dv, float h, cudaStream_t st) {
    if (g.cols % 4 == 0)
        k_divergence_v4<<<grid_v4(g), dim3(kVx, kVy), 0, st>>>(g, u, v, dv, -0.5f * h);
    else
        k_divergence<<<grid2d(g), dim3(kBx, kBy), 0, st>>>(g, u, v, dv, -0.5f * h);
}

// ------------------------------------------------------------------------------- gradient
// remove_p_kernel (gpu.cu:193-206) + set_boundary_opposite_horizontal(u) / _vertical(v)
// (gpu.cu:402-403).  Out of place so that edge threads can re-evaluate their inward neighbour.
__global__ void __launch_bounds__(kBx* kBy) k_gradient(Geom g, const float* __restrict__ p,
                                                      const float* __restrict__ u_in,
                                                      const float* __restrict__ v_in, float* __restrict__ u_out,
                                                      float* __restrict__ v_out, float h) {
    F2D_CELL_IJ();
    // u and v have different boundary kinds but the same source cell
    const CellSrc cu = classify_cell(g, i, j, F2D_BND_OPPOSITE_HORIZONTAL);
    const CellSrc cv = classify_cell(g, i, j, F2D_BND_OPPOSITE_VERTICAL);
    const size_t o = (size_t)i * g.pitch + j;
    if (cu.cls == CELL_KEEP) {
        u_out[o] = u_in[o];
        v_out[o] = v_in[o];
        return;
    }
    const size_t so = (size_t)cu.si * g.pitch + cu.sj;
    const float un = gradient_update(u_in[so], p[so + 1], p[so - 1], h);
    const float vn = gradient_update(v_in[so], p[so + g.pitch], p[so - g.pitch], h);
    u_out[o] = apply_sign(un, cu.negate);
    v_out[o] = apply_sign(vn, cv.negate);
}

__global__ void __launch_bounds__(kVx* kVy) k_gradient_v4(Geom g, const float* __restrict__ p,
                                                         const float* __restrict__ u_in,
                                                         const float* __restrict__ v_in, float* __restrict__ u_out,
                                                         float* __restrict__ v_out, float h) {
    F2D_ROW_J4();
    const RowSrc rs = classify_row(g, i);
    const size_t o = (size_t)i * g.pitch + j;
    const float4 u_here = ld4(u_in + o), v_here = ld4(v_in + o);
    if (rs.keep) {
        st4(u_out + o, u_here);
        st4(v_out + o, v_here);
        return;
    }
    const size_t so = (size_t)rs.si * g.pitch + j;
    const float4 pc = ld4(p + so), ps = ld4(p + so + g.pitch), pn = ld4(p + so - g.pitch);
    const float pl = (j > 0) ? __ldg(p + so - 1) : 0.f;
    const float pr = (j + 4 < g.cols) ? __ldg(p + so + 4) : 0.f;
    const float4 us = rs.edge_row ? ld4(u_in + so) : u_here, vsrc = rs.edge_row ? ld4(v_in + so) : v_here;
    float4 un, vn;
    un.x = gradient_update(us.x, pc.y, pl, h);
    un.y = gradient_update(us.y, pc.z, pc.x, h);
    un.z = gradient_update(us.z, pc.w, pc.y, h);
    un.w = gradient_update(us.w, pr, pc.z, h);
    vn.x = gradient_update(vsrc.x, ps.x, pn.x, h);
    vn.y = gradient_update(vsrc.y, ps.y, pn.y, h);
    vn.z = gradient_update(vsrc.z, ps.z, pn.z, h);
    vn.w = gradient_update(vsrc.w, ps.w, pn.w, h);
    // edge rows: u copies (opposite_horizontal), v negates (opposite_vertical) the adjacent interior row
    if (rs.edge_row) vn = neg4(vn, true);
    // edge columns of interior rows: u negates, v copies its neighbour; corners keep the input
    if (j == 0) {
        un.x = rs.edge_row ? u_here.x : -un.y;
        vn.x = rs.edge_row ? v_here.x : vn.y;
    }
    if (j + 3 == g.cols - 1) {
        un.w = rs.edge_row ? u_here.w : -un.z;
        vn.w = rs.edge_row ? v_here.w : vn.z;
    }
    st4(u_out + o, un);
    st4(v_out + o, vn);
}

void launch_gradient(const Geom& g, const float* p, const float* u_in, const float* v_in, float* u_out,
                     float* v_out, float h, cudaStream_t st) {
    if (g.cols % 4 == 0)
        k_gradient_v4<<<grid_v4(g), dim3(kVx, kVy), 0, st>>>(g, p, u_in, v_in, u_out, v_out, h);
    else
        k_gradient<<<grid2d(g), dim3(kBx, kBy), 0, st>>>(g, p, u_in, v_in, u_out, v_out, h);
}

// ------------------------------------------------------------------------ advect (gather)
// advect_kernel (gpu.cu:99-129) for u AND v in one pass: both are advected by the same (U0,V0)
// (gpu.cu:248-251), so the back-traced point and the bilinear weights are shared.  The gathers go
// through the read-only path; with CFL-bounded displacements they hit L1/L2.  Global row indices
// are used for y so that a row slab rounds exactly like the single-GPU run.
__global__ void __launch_bounds__(kBx* kBy) k_advect_velocity(Geom g, const float* __restrict__ u0,
                                                             const float* __restrict__ v0,
                                                             float* __restrict__ u_out, float* __restrict__ v_out,
                                                             float dt0, int own_begin, int own_end, int valid_lo, int valid_hi,
                                                             int* oob_flag) {
    F2D_CELL_IJ();
    const CellSrc cu = classify_cell(g, i, j, F2D_BND_OPPOSITE_HORIZONTAL);
    const CellSrc cv = classify_cell(g, i, j, F2D_BND_OPPOSITE_VERTICAL);
    const size_t o = (size_t)i * g.pitch + j;
    if (cu.cls == CELL_KEEP) {
        u_out[o] = u0[o];
        v_out[o] = v0[o];
        return;
    }
    const size_t so = (size_t)cu.si * g.pitch + cu.sj;
    float x = __fmaf_rn(-__ldg(u0 + so), dt0, (float)cu.sj);
    float y = __fmaf_rn(-__ldg(v0 + so), dt0, (float)(g.grow0 + cu.si));
    x = fmaxf(1.5f, fminf((float)g.cols - 1.5f, x));
    y = fmaxf(1.5f, fminf((float)g.grows - 1.5f, y));
    const Bilinear b = bilinear_setup(x, y);
    // local rows li0, li0 + 1 of the gather.  [valid_lo, valid_hi) are the rows of u0 / v0 known to be valid (the halo
    // rows refreshed by the last exchange minus what later stencils invalidated): an owned cell whose back-trace leaves
    // them broke the displacement bound the exchange schedule was built for -> error flag, never a silent stale read
    int li0 = b.i0 - g.grow0;
    if ((li0 < valid_lo || li0 + 1 >= valid_hi) && i >= own_begin && i < own_end) *oob_flag = 1;
    li0 = max(0, min(g.rows - 2, li0));
    const size_t a = (size_t)li0 * g.pitch + b.j0;
    const float un = bilinear_gather(b, __ldg(u0 + a), __ldg(u0 + a + 1), __ldg(u0 + a + g.pitch), __ldg(u0 + a + g.pitch + 1));
    const float vn = bilinear_gather(b, __ldg(v0 + a), __ldg(v0 + a + 1), __ldg(v0 + a + g.pitch), __ldg(v0 + a + g.pitch + 1));
    u_out[o] = apply_sign(un, cu.negate);
    v_out[o] = apply_sign(vn, cv.negate);
}

// float4 version (cols % 4 == 0): one thread = four consecutive cells of one row.  The cells' own velocities (the
// back-trace) are one 16-byte load per component -- of the adjacent interior row for a global edge row, and the edge
// columns take their neighbour component of the same float4 (classify_cell's (si, sj)) --, the results two 16-byte
// stores; the 2 x 4 bilinear taps per cell stay scalar read-only loads (they hit L1/L2 for CFL-bounded steps).
__global__ void __launch_bounds__(kVx* kVy) k_advect_velocity_v4(Geom g, const float* __restrict__ u0,
                                                                const float* __restrict__ v0, float* __restrict__ u_out,
                                                                float* __restrict__ v_out, float dt0, int own_begin, int own_end,
                                                                int valid_lo, int valid_hi, int* oob_flag) {
    F2D_ROW_J4();
    const size_t o = (size_t)i * g.pitch + j;
    const RowSrc rs = classify_row(g, i);
    if (rs.keep) {
        st4(u_out + o, ld4(u0 + o));
        st4(v_out + o, ld4(v0 + o));
        return;
    }
    const size_t so = (size_t)rs.si * g.pitch + j;
    const float4 us = ld4(u0 + so), vs = ld4(v0 + so);
    const float uc[4] = {us.x, us.y, us.z, us.w}, vc[4] = {vs.x, vs.y, vs.z, vs.w};
    const bool first = (j == 0), last = (j + 3 == g.cols - 1);
    const bool own = (i >= own_begin && i < own_end);
    const float ytrace = (float)(g.grow0 + rs.si);
    float ur[4], vr[4];
    bool oob = false;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        // source column of cell c: the edge columns evaluate their inward neighbour (gpu.cu:16-17, 31-32)
        const int cs = (first && c == 0) ? 1 : ((last && c == 3) ? 2 : c);
        float x = __fmaf_rn(-uc[cs], dt0, (float)(j + cs));
        float y = __fmaf_rn(-vc[cs], dt0, ytrace);
        x = fmaxf(1.5f, fminf((float)g.cols - 1.5f, x));
        y = fmaxf(1.5f, fminf((float)g.grows - 1.5f, y));
        const Bilinear b = bilinear_setup(x, y);
        int li0 = b.i0 - g.grow0;
        oob |= (li0 < valid_lo || li0 + 1 >= valid_hi);
        li0 = max(0, min(g.rows - 2, li0));
        const size_t a = (size_t)li0 * g.pitch + b.j0;
        ur[c] = bilinear_gather(b, __ldg(u0 + a), __ldg(u0 + a + 1), __ldg(u0 + a + g.pitch), __ldg(u0 + a + g.pitch + 1));
        vr[c] = bilinear_gather(b, __ldg(v0 + a), __ldg(v0 + a + 1), __ldg(v0 + a + g.pitch), __ldg(v0 + a + g.pitch + 1));
    }
    if (oob && own) *oob_flag = 1;
    // signs of the fused boundary pass: u flips in the edge columns, v in the edge rows (gpu.cu:26-54)
    if (first) ur[0] = -ur[0];
    if (last) ur[3] = -ur[3];
    if (rs.edge_row) {
#pragma unroll
        for (int c = 0; c < 4; ++c) vr[c] = -vr[c];
        // corners are never written by the reference (gpu.cu:15-23): pass the input through
        if (first) {
            ur[0] = u0[o];
            vr[0] = v0[o];
        }
        if (last) {
            ur[3] = u0[o + 3];
            vr[3] = v0[o + 3];
        }
    }
    st4(u_out + o, make_float4(ur[0], ur[1], ur[2], ur[3]));
    st4(v_out + o, make_float4(vr[0], vr[1], vr[2], vr[3]));
}

void launch_advect_velocity(const Geom& g, const float* u0, const float* v0, float* u_out, float* v_out,
                            float dt0, int own_begin, int own_end, int valid_lo, int valid_hi, int* oob_flag, cudaStream_t st) {
    if (g.cols % 4 == 0)
        k_advect_velocity_v4<<<grid_v4(g), dim3(kVx, kVy), 0, st>>>(g, u0, v0, u_out, v_out, dt0, own_begin, own_end, valid_lo, valid_hi, oob_flag);
    else
        k_advect_velocity<<<grid2d(g), dim3(kBx, kBy), 0, st>>>(g, u0, v0, u_out, v_out, dt0, own_begin, own_end, valid_lo, valid_hi, oob_flag);
}

// rows [row0, row0 + nrows) of f += src (src is a dense nrows x pitch block): the receiving end of the
// reverse halo exchange of the density scatter (multi-GPU)
__global__ void k_add_rows(float* __restrict__ f, const float* __restrict__ src, size_t n4) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    float4 a = reinterpret_cast<float4*>(f)[i];
    const float4 b = reinterpret_cast<const float4*>(src)[i];
    a.x = __fadd_rn(a.x, b.x);
    a.y = __fadd_rn(a.y, b.y);
    a.z = __fadd_rn(a.z, b.z);
    a.w = __fadd_rn(a.w, b.w);
    reinterpret_cast<float4*>(f)[i] = a;
}

void launch_add_rows(const Geom& g, float* f, int row0, int nrows, const float* src, cudaStream_t st) {
    const size_t n4 = (size_t)nrows * g.pitch / 4;
    k_add_rows<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(f + (size_t)row0 * g.pitch, src, n4);
}

// ----------------------------------------------------------------------- advect (scatter)
// advect_trace_kernel (gpu.cu:131-162): forward trace of every interior cell and a bilinear splat
// with four float atomics (RED.ADD.F32, flush-to-zero like the reference's).  `out` is zeroed by
// the caller (gpu.cu:337).  The boundary pass of gpu.cu:355 is fused into launch_smooth_bnd.
__global__ void __launch_bounds__(kBx* kBy) k_scatter_density(Geom g, const float* __restrict__ src,
                                                             const float* __restrict__ u,
                                                             const float* __restrict__ v, float* out, float dt0,
                                                             int own_begin, int own_end, int* oob_flag) {
    F2D_CELL_IJ();
    const int gi = g.grow0 + i;
    if (i < own_begin || i >= own_end || gi < 1 || gi > g.grows - 2 || j < 1 || j > g.cols - 2) return;
    const size_t o = (size_t)i * g.pitch + j;
    const float x = __fmaf_rn(__ldg(u + o), dt0, (float)j);
    const float y = __fmaf_rn(__ldg(v + o), dt0, (float)gi);
    if (x < 0.5f || x > (float)g.cols - 1.5f || y < 0.5f || y > (float)g.grows - 1.5f) return;
    const Bilinear b = bilinear_setup(x, y);
    const int li0 = b.i0 - g.grow0;
    if (li0 < 0 || li0 + 1 >= g.rows) {  // displacement exceeded the slab halo: report, never fault
        *oob_flag = 1;
        return;
    }
    const float val = __ldg(src + o);
    float* t = out + (size_t)li0 * g.pitch + b.j0;
    atomicAdd(t, __fmul_rn(__fmul_rn(b.s1, b.s3), val));
    atomicAdd(t + g.pitch, __fmul_rn(__fmul_rn(b.s1, b.s2), val));
    atomicAdd(t + 1, __fmul_rn(__fmul_rn(b.s0, b.s3), val));
    atomicAdd(t + g.pitch + 1, __fmul_rn(__fmul_rn(b.s0, b.s2), val));
}

void launch_scatter_density(const Geom& g, const float* src, const float* u, const float* v, float* out,
                            float dt0, int own_begin, int own_end, int* oob_flag, cudaStream_t st) {
    // one thread per source cell on purpose: the 32 lanes of a warp then hit consecutive addresses with each of the four
    // RED.ADD (one or two 128-byte lines per instruction).  A float4-per-thread version (three 16-byte loads, 16 REDs per
    // thread) was measured at 174 us against 96 us at 4096^2: its REDs touch four lines each and the thread serialises 16
    // of them (lg_throttle 38 cycles per instruction, profiles/ncu_advect_scatter_r02.md)
    k_scatter_density<<<grid2d(g), dim3(kBx, kBy), 0, st>>>(g, src, u, v, out, dt0, own_begin, own_end, oob_flag);
}

// ------------------------------------------------------------------- smooth + boundary pass
// set_boundary_continuous after the scatter (gpu.cu:355) fused with smooth_kernel (gpu.cu:87-97,
// no boundary pass afterwards, gpu.cu:314-323).  B(in) denotes `in` after the boundary pass:
// edge cells read their inward neighbour, everything else reads itself.
__device__ __forceinline__ float bnd_read(const Geom& g, const float* __restrict__ in, int i, int j) {
    const CellSrc c = classify_cell(g, i, j, F2D_BND_CONTINUOUS);
    return __ldg(in + (size_t)c.si * g.pitch + c.sj);  // CELL_KEEP has (si,sj) == (i,j)
}

template <bool SMOOTH>
__global__ void __launch_bounds__(kBx* kBy) k_smooth_bnd(Geom g, const float* __restrict__ in, float* __restrict__ out) {
    F2D_CELL_IJ();
    const int gi = g.grow0 + i;
    const size_t o = (size_t)i * g.pitch + j;
    const bool interior = gi >= 1 && gi <= g.grows - 2 && j >= 1 && j <= g.cols - 2 && i >= 1 && i <= g.rows - 2;
    if (!SMOOTH || !interior) {
        out[o] = bnd_read(g, in, i, j);
        return;
    }
    out[o] = smooth_update(__ldg(in + o), bnd_read(g, in, i, j - 1), bnd_read(g, in, i, j + 1),
                           bnd_read(g, in, i - 1, j), bnd_read(g, in, i + 1, j));
}

// float4 version (cols % 4 == 0).  B(in) at the neighbours of an interior row: the row above row 1 / below
// row N-2 and the column left of column 1 / right of column N-2 are edge cells whose boundary value is the
// interior cell itself.
template <bool SMOOTH>
__global__ void __launch_bounds__(kVx* kVy) k_smooth_bnd_v4(Geom g, const float* __restrict__ in, float* __restrict__ out) {
    F2D_ROW_J4();
    const int gi = g.grow0 + i;
    const size_t o = (size_t)i * g.pitch + j;
    const bool top = (gi == 0), bottom = (gi == g.grows - 1);
    const bool first = (j == 0), last = (j + 3 == g.cols - 1);
    if ((i == 0 && !top) || (i == g.rows - 1 && !bottom)) {  // slab-local edge row: pass through
        st4(out + o, ld4(in + o));
        return;
    }
    if (top || bottom) {  // global edge row: copy of the adjacent interior row, corners keep their value
        const float4 here = ld4(in + o);
        float4 r = ld4(in + (top ? o + g.pitch : o - g.pitch));
        if (first) r.x = here.x;
        if (last) r.w = here.w;
        st4(out + o, r);
        return;
    }
    const float4 c = ld4(in + o);
    float4 r = c;
    if (SMOOTH) {
        const bool n_edge = (gi - 1 == 0), s_edge = (gi + 1 == g.grows - 1);
        const bool can = (i >= 1 && i <= g.rows - 2);  // neighbours exist locally (always true for global interior rows of a full grid)
        if (can) {
            const float4 n = n_edge ? c : ld4(in + o - g.pitch);
            const float4 sN = s_edge ? c : ld4(in + o + g.pitch);
            const float wl = first ? 0.f : __ldg(in + o - 1);
            const float er = last ? 0.f : __ldg(in + o + 4);
            // B(in)(i, 0) == in(i, 1) and B(in)(i, N-1) == in(i, N-2)
            r.x = smooth_update(c.x, wl, c.y, n.x, sN.x);
            r.y = smooth_update(c.y, first ? c.y : c.x, c.z, n.y, sN.y);
            r.z = smooth_update(c.z, c.y, last ? c.z : c.w, n.z, sN.z);
            r.w = smooth_update(c.w, c.z, er, n.w, sN.w);
        }
    }
    // edge columns of an interior row: boundary value = the neighbouring interior cell of the INPUT
    if (first) r.x = c.y;
    if (last) r.w = c.z;
    st4(out + o, r);
}

void launch_smooth_bnd(const Geom& g, const float* in, float* out, bool do_smooth, cudaStream_t st) {
    if (g.cols % 4 == 0 && g.cols >= 8) {
        if (do_smooth)
            k_smooth_bnd_v4<true><<<grid_v4(g), dim3(kVx, kVy), 0, st>>>(g, in, out);
        else
            k_smooth_bnd_v4<false><<<grid_v4(g), dim3(kVx, kVy), 0, st>>>(g, in, out);
        return;
    }
    if (do_smooth)
        k_smooth_bnd<true><<<grid2d(g), dim3(kBx, kBy), 0, st>>>(g, in, out);
    else
        k_smooth_bnd<false><<<grid2d(g), dim3(kBx, kBy), 0, st>>>(g, in, out);
}

// smooth_kernel exactly as a stand-alone stage (gpu.cu:87-97): interior only, edges untouched.
__global__ void __launch_bounds__(kBx* kBy) k_smooth_plain(Geom g, const float* __restrict__ in, float* __restrict__ out) {
    F2D_CELL_IJ();
    const int gi = g.grow0 + i;
    if (gi < 1 || gi > g.grows - 2 || j < 1 || j > g.cols - 2 || i < 1 || i > g.rows - 2) return;
    const size_t o = (size_t)i * g.pitch + j;
    out[o] = smooth_update(in[o], in[o - 1], in[o + 1], in[o - g.pitch], in[o + g.pitch]);
}

void launch_smooth_plain(const Geom& g, const float* in, float* out, cudaStream_t st) {
    k_smooth_plain<<<grid2d(g), dim3(kBx, kBy), 0, st>>>(g, in, out);
}

// --------------------------------------------------------------------- in-place set_bnd
// set_boundary_*_kernel (gpu.cu:11-54) as a stand-alone stage (parity tests, f2d_stage_set_bnd).
__global__ void k_set_bnd_inplace(Geom g, float* f, int kind) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    const bool neg_c = (kind == F2D_BND_OPPOSITE_HORIZONTAL), neg_r = (kind == F2D_BND_OPPOSITE_VERTICAL);
    if (m < g.rows) {  // left / right columns of every local row that is a global interior row
        const int gi = g.grow0 + m;
        if (gi >= 1 && gi <= g.grows - 2) {
            float* r = f + (size_t)m * g.pitch;
            r[0] = apply_sign(r[1], neg_c);
            r[g.cols - 1] = apply_sign(r[g.cols - 2], neg_c);
        }
    }
    if (m >= 1 && m <= g.cols - 2) {  // global top / bottom rows, corners excluded
        if (g.grow0 == 0) f[m] = apply_sign(f[g.pitch + m], neg_r);
        if (g.grow0 + g.rows == g.grows)
            f[(size_t)(g.rows - 1) * g.pitch + m] = apply_sign(f[(size_t)(g.rows - 2) * g.pitch + m], neg_r);
    }
}

void launch_set_bnd_inplace(const Geom& g, float* f, int kind, cudaStream_t st) {
    const int n = max(g.rows, g.cols);
    k_set_bnd_inplace<<<(n + 127) / 128, 128, 0, st>>>(g, f, kind);
}

}  // namespace f2d
