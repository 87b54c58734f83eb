// f2d_gs.cu -- the fluid_solver_cpu-compatible stages (F2D_SEM_CPU) for sm_100a.
//
// fluid_solver_cpu (src/fluid_solver_cpu.cpp) differs from fluid_solver_gpu in arithmetic, not in physics:
// in-place Gauss-Seidel relaxations (:104-113, :196-204), no FMA contraction (g++ -O2, generic x86-64), float
// divides, corners averaged by set_boundary (:44-47), a density scatter that accumulates in lexicographic
// source order (:127-152) and no smooth.  This file reproduces those bits on the GPU:
//   k_gs_relax        all sweeps of a relaxation as ONE wavefront of 32x32 tiles over the in-place array
//                     (dependencies, tile core and edge ownership: f2d_gs_tile.h), one warp per (sweep, row band),
//                     bands chained through progress counters in global memory;
//   k_scatter_ordered the density scatter as a gather: every target cell scans the sources that can reach it in
//                     lexicographic order, so the float additions happen in the CPU's order (no atomics);
//   k_scatter_keys    every source's landing cell and the reach of that scan (largest displacement, in cells);
//   k_add_sources_nofma, k_advect_velocity_nofma, k_corners_avg.
// Divergence and gradient subtract have the same operations in both reference solvers
// (cpp:190-193, :207-211 vs gpu.cu:173-174, :202-203) and reuse the kernels of f2d_kernels_simple.cu.
#include "f2d_gs_tile.h"
#include "f2d_kernels.cuh"
#include "f2d_scatter_core.h"

#ifdef F2D_GS_TIMING
// Profiling build only (tools/gs_timing.py): cycles per phase of a tile, summed over all warps.
__device__ unsigned long long g_gs_cycles[8];
#define GS_T(var) const long long var = clock64()
#define GS_ACC(i, a, b) \
    if (lane == 0 && c > 0) atomicAdd(&g_gs_cycles[i], (unsigned long long)((b) - (a)))
extern "C" __attribute__((visibility("default"))) int f2d_debug_gs_cycles(unsigned long long* out8, int reset) {
    if (cudaMemcpyFromSymbol(out8, g_gs_cycles, sizeof(g_gs_cycles)) != cudaSuccess) return 1;
    if (reset) {
        unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (cudaMemcpyToSymbol(g_gs_cycles, z, sizeof(z)) != cudaSuccess) return 1;
    }
    return 0;
}
#else
#define GS_T(var)
#define GS_ACC(i, a, b)
#endif

#ifndef F2D_GS_MINB
#define F2D_GS_MINB 4  // resident CTAs per SM the register allocator must allow (A/B: tools/build_variants.sh gs)
#endif

namespace f2d {

namespace {
constexpr int kGsWarps = 4;
constexpr unsigned long long kWaitLimitNs = 30ull * 1000ull * 1000ull * 1000ull;  // a wait longer than this raises the error flag

__device__ __forceinline__ unsigned ld_relaxed(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void st_release(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// The whole warp, convergently: block until *flag >= need.  Every lane reads the same word (one transaction) and
// takes lane 0's value, so control flow stays warp-uniform -- a poll loop run by lane 0 alone leaves the warp split
// in two for the rest of the tile (measured: every later __syncwarp took the divergent slow path and the step loop
// issued twice, 46k instead of 6k cycles per tile).  Bounded: a wait that never ends (a scheduling assumption
// broken) raises *err and returns, and every other wait then returns at once, so the launch always terminates.
__device__ __forceinline__ unsigned wait_at_least(const unsigned* flag, unsigned need, int* err, int lane) {
    // polls are relaxed loads (served by L2, no L1 invalidation per poll); one acquire fence once the value is in
    unsigned v = __shfl_sync(0xffffffffu, ld_relaxed(flag), 0);
    unsigned spins = 0, ns = 0;
    unsigned long long t_start = 0;
    while (v < need) {
        // a neighbour one tile behind arrives within a few polls: spin, then back off gently.  One that is several
        // tiles behind (a band that has not reached this column yet, a later sweep waiting for its turn) cannot
        // arrive sooner than a tile takes (~4 us): sleep accordingly and keep the memory system free for the
        // warps that are on the wavefront.
        const unsigned behind = need - v;
        if (behind > 1u)
            __nanosleep(behind > 8u ? 16000u : 2000u * (behind - 1u));
        else if (ns)
            __nanosleep(ns);
        if (spins >= 8 && ns < 256) ns = ns ? 2 * ns : 32;
        v = __shfl_sync(0xffffffffu, ld_relaxed(flag), 0);
        if ((++spins & 255u) == 0u) {
            if (__shfl_sync(0xffffffffu, *reinterpret_cast<volatile int*>(err), 0)) break;
            const unsigned long long now = __shfl_sync(0xffffffffu, global_ns(), 0);
            if (t_start == 0) t_start = now;
            if (now - t_start > kWaitLimitNs) {
                if (lane == 0) atomicExch(err, 1);
                break;
            }
        }
    }
    asm volatile("fence.acq_rel.gpu;" ::: "memory");  // acquire: the tile data behind the counter is read after this point
    return v;
}
}  // namespace

// One warp = one (problem, sweep, row band); it marches over the column tiles of its band.  Warps take their
// work from a ticket counter in the order (sweep, band, problem), so every dependency of a warp points to a
// LOWER ticket, i.e. to a warp that is already running or finished: the wavefront cannot deadlock however many
// warps the grid has and however few are resident.
template <bool DIFFUSE>
__global__ void __launch_bounds__(kGsWarps * 32, F2D_GS_MINB) k_gs_relax(GsBatch b) {
    __shared__ float smem[kGsWarps][gs::kWarpFloats];
    const int lane = threadIdx.x & 31;
    float* tile = smem[threadIdx.x >> 5] + gs::kPad;
    float* rt = tile + gs::kTileFloats;

    unsigned ticket = 0;
    if (lane == 0) ticket = atomicAdd(b.ticket, 1u);
    ticket = __shfl_sync(0xffffffffu, ticket, 0);
    const gs::Shape s = gs::make_shape(b.rows, b.cols, b.pitch);
    if (ticket >= (unsigned)b.n * (unsigned)b.sweeps * (unsigned)s.nb) return;
    const int f = (int)(ticket % (unsigned)b.n);
    const unsigned rest = ticket / (unsigned)b.n;
    const int w = (int)(rest % (unsigned)s.nb), k = (int)(rest / (unsigned)s.nb);

    float* x = b.p[f].x;
    const float* rhs = b.p[f].rhs;
    const float a = b.p[f].a, cdiv = b.p[f].c;
    const int kind = b.p[f].kind;
    unsigned* done = b.flags + (size_t)f * b.sweeps * s.nb;
    unsigned seen[3] = {0u, 0u, 0u};  // last value read from each dependency's counter (they only grow)

    float pf[gs::kPf];  // the next tile's interior + right-hand side, in flight while the current tile computes
    // c = -1 is a lead-in that only waits for tile 0's neighbours and requests tile 0; tile c then commits what
    // was requested one iteration earlier, requests tile c + 1 and computes (one request site, one commit site)
    for (int c = -1; c < s.nt; ++c) {
        GS_T(t0);
        {
            const gs::Deps d = gs::tile_deps(s, k, w, c < 0 ? 0 : c);  // warp-uniform
#pragma unroll
            for (int i = 0; i < 3; ++i)
                if (i < d.n && seen[i] < d.need[i]) seen[i] = wait_at_least(done + d.idx[i], d.need[i], b.err, lane);
        }
        GS_T(t1);
        gs::Tile t = gs::make_tile(s, w, c < 0 ? 0 : c);
        if (c >= 0) {
            const gs::Frame fr = gs::tile_frame_load(s, t, x, tile, lane, c == 0);
            __syncwarp();  // every lane has read the previous tile's last column
            gs::tile_commit(t, pf, tile, rt, lane);
            gs::tile_frame_store(t, fr, tile, lane);
            __syncwarp();
        }
        GS_T(t2);
        if (c + 1 < s.nt) gs::tile_prefetch(s, gs::make_tile(s, w, c + 1), x, rhs, pf, lane);
        if (c < 0) continue;
        GS_T(t3);
        gs::StepRegs g;
        gs::tile_step_init(t, tile, rt, lane, g);
        float north = 0.f;
        const int nsteps = t.nr + t.nc - 1;
        for (int step = 0; step < nsteps; ++step) {
            const float v = gs::tile_step<DIFFUSE>(t, lane, step, a, cdiv, g, north);
            north = __shfl_up_sync(0xffffffffu, v, 1);  // lane l's north neighbour of the next step
            __syncwarp();
        }
        GS_T(t4);
        gs::tile_store(s, t, kind, x, tile, lane);
        __syncwarp();  // orders every lane's stores before lane 0's release (fence + store), cumulatively
        GS_T(t5);
        if (lane == 0) st_release(done + (size_t)k * s.nb + w, (unsigned)(c + 1));
        GS_T(t6);
        GS_ACC(0, t0, t1);
        GS_ACC(1, t1, t2);
        GS_ACC(2, t2, t3);
        GS_ACC(3, t3, t4);
        GS_ACC(4, t4, t5);
        GS_ACC(5, t5, t6);
        GS_ACC(6, 0, 1);
    }
}

size_t gs_flag_words(int rows, int nproblems, int sweeps) {
    const int nb = (rows - 2 + gs::kBand - 1) / gs::kBand;
    return (size_t)nproblems * (size_t)sweeps * (size_t)nb + 1;  // + the ticket counter
}

void launch_gs_relax(const GsBatch& b, bool diffuse, cudaStream_t st) {
    const gs::Shape s = gs::make_shape(b.rows, b.cols, b.pitch);
    const size_t warps = (size_t)b.n * b.sweeps * s.nb;
    const unsigned ctas = (unsigned)((warps + kGsWarps - 1) / kGsWarps);
    if (diffuse)
        k_gs_relax<true><<<ctas, kGsWarps * 32, 0, st>>>(b);
    else
        k_gs_relax<false><<<ctas, kGsWarps * 32, 0, st>>>(b);
}

// ------------------------------------------------------------------------------------ corners
// The tail of fluid_solver_cpu::set_boundary_* (cpp:44-47, :61-64, :79-82): each corner is the mean of its two
// edge neighbours.  Runs after a kernel that produced the edges.
__global__ void k_corners_avg(CornerBatch b, int rows, int cols, int pitch) {
    const int t = threadIdx.x;
    if (t >= 4 * b.n) return;
    float* f = b.f[t >> 2];
    const int ci = (t & 2) ? rows - 1 : 0, cj = (t & 1) ? cols - 1 : 0;
    const int ni = (t & 2) ? rows - 2 : 1, nj = (t & 1) ? cols - 2 : 1;
    // 0.5f * (f(ci, nj) + f(ni, cj)): the row neighbour first, as the reference writes it
    f[(size_t)ci * pitch + cj] = __fmul_rn(0.5f, __fadd_rn(f[(size_t)ci * pitch + nj], f[(size_t)ni * pitch + cj]));
}

void launch_corners_avg(const Geom& g, const CornerBatch& b, cudaStream_t st) {
    k_corners_avg<<<1, 32, 0, st>>>(b, g.rows, g.cols, g.pitch);
}

// -------------------------------------------------------------------------------- add_sources
// cpp:85-93: f += dt * s with the product rounded first (no FMA); interior only, other cells pass through.
__global__ void __launch_bounds__(256) k_add_sources_nofma(Geom g, AddSourceBatch b, float dt) {
    const int j = blockIdx.x * 32 + threadIdx.x, i = blockIdx.y * 8 + threadIdx.y;
    if (i >= g.rows || j >= g.cols) return;
    const size_t o = (size_t)i * g.pitch + j;
    const float* f = b.f[blockIdx.z];
    float* out = b.o[blockIdx.z];
    const bool interior = i >= 1 && i <= g.rows - 2 && j >= 1 && j <= g.cols - 2;
    if (interior)
        out[o] = __fadd_rn(f[o], __fmul_rn(dt, __ldg(b.s[blockIdx.z] + o)));
    else if (out != f)
        out[o] = f[o];
}

void launch_add_sources_nofma(const Geom& g, const AddSourceBatch& b, float dt, cudaStream_t st) {
    k_add_sources_nofma<<<dim3((g.cols + 31) / 32, (g.rows + 7) / 8, b.n), dim3(32, 8), 0, st>>>(g, b, dt);
}

// ------------------------------------------------------------------------------ advect (gather)
// cpp:153-174 for u and v in one pass (both are advected by the same copies, cpp:26-29): separate multiply and
// subtract for the back-trace, and the bilinear form s3*(s1*a00 + s0*a01) + s2*(s1*a10 + s0*a11) without FMA.
// Edges are produced by evaluating the inward neighbour (cpp:176), corners keep the input until k_corners_avg.
__device__ __forceinline__ float bilinear_nofma(const Bilinear& b, float a00, float a01, float a10, float a11) {
    const float top = __fadd_rn(__fmul_rn(b.s1, a00), __fmul_rn(b.s0, a01));
    const float bot = __fadd_rn(__fmul_rn(b.s1, a10), __fmul_rn(b.s0, a11));
    return __fadd_rn(__fmul_rn(b.s3, top), __fmul_rn(b.s2, bot));
}

__global__ void __launch_bounds__(256) k_advect_velocity_nofma(Geom g, const float* __restrict__ u0,
                                                              const float* __restrict__ v0, float* __restrict__ u_out,
                                                              float* __restrict__ v_out, float dt0) {
    const int j = blockIdx.x * 32 + threadIdx.x, i = blockIdx.y * 8 + threadIdx.y;
    if (i >= g.rows || j >= g.cols) return;
    const CellSrc cu = classify_cell(g, i, j, F2D_BND_OPPOSITE_HORIZONTAL);
    const CellSrc cv = classify_cell(g, i, j, F2D_BND_OPPOSITE_VERTICAL);
    const size_t o = (size_t)i * g.pitch + j;
    if (cu.cls == CELL_KEEP) {
        u_out[o] = u0[o];
        v_out[o] = v0[o];
        return;
    }
    const size_t so = (size_t)cu.si * g.pitch + cu.sj;
    float x = __fsub_rn((float)cu.sj, __fmul_rn(dt0, __ldg(u0 + so)));
    float y = __fsub_rn((float)cu.si, __fmul_rn(dt0, __ldg(v0 + so)));
    x = fmaxf(1.5f, fminf((float)g.cols - 1.5f, x));
    y = fmaxf(1.5f, fminf((float)g.rows - 1.5f, y));
    const Bilinear b = bilinear_setup(x, y);
    const size_t a = (size_t)b.i0 * g.pitch + b.j0;
    const float un = bilinear_nofma(b, __ldg(u0 + a), __ldg(u0 + a + 1), __ldg(u0 + a + g.pitch), __ldg(u0 + a + g.pitch + 1));
    const float vn = bilinear_nofma(b, __ldg(v0 + a), __ldg(v0 + a + 1), __ldg(v0 + a + g.pitch), __ldg(v0 + a + g.pitch + 1));
    u_out[o] = apply_sign(un, cu.negate);
    v_out[o] = apply_sign(vn, cv.negate);
}

void launch_advect_velocity_nofma(const Geom& g, const float* u0, const float* v0, float* u_out, float* v_out, float dt0,
                                  cudaStream_t st) {
    k_advect_velocity_nofma<<<dim3((g.cols + 31) / 32, (g.rows + 7) / 8), dim3(32, 8), 0, st>>>(g, u0, v0, u_out, v_out, dt0);
}

// ----------------------------------------------------------------------- advect (ordered scatter)
// cpp:127-152 turned inside out (algorithm and per-cell functions: f2d_scatter_core.h).  Two passes, no atomics on
// the data:
//   k_scatter_keys     one thread per source: its landing cell as a linear index (or kNoKey), and the largest
//                      displacement max(|dt0*u|, |dt0*v|) over the sources that land inside the grid, as the bit pattern
//                      of a non-negative float (integer max == float max there; a NaN compares above everything and
//                      widens the scan to the whole grid);
//   k_scatter_ordered  one thread per interior target: visits, in lexicographic order, every source whose displacement
//                      can reach it (|di|, |dj| <= R = ceil(max displacement) + 1) and adds the share of those that
//                      land on it.  Edge cells and corners are overwritten by the boundary pass that follows
//                      (cpp:176), so only the interior is produced.
__global__ void __launch_bounds__(256) k_scatter_keys(Geom g, const float* __restrict__ u, const float* __restrict__ v, float dt0,
                                                     unsigned* __restrict__ keys, unsigned* disp_bits) {
    const int j = blockIdx.x * 32 + threadIdx.x, i = blockIdx.y * 8 + threadIdx.y;
    unsigned m = 0u;
    if (i < g.rows && j < g.cols) {
        const size_t o = (size_t)i * g.pitch + j;
        unsigned key = sc::kNoKey;
        if (i >= 1 && i <= g.rows - 2 && j >= 1 && j <= g.cols - 2) {
            const float uu = __ldg(u + o), vv = __ldg(v + o);
            key = sc::source_key(g.rows, g.cols, g.pitch, i, j, uu, vv, dt0);
            // only sources that land somewhere widen the scan: a skipped source (cpp:134) reaches no target, however
            // far it would have travelled
            if (key != sc::kNoKey) m = max(__float_as_uint(fabsf(__fmul_rn(dt0, uu))), __float_as_uint(fabsf(__fmul_rn(dt0, vv))));
        }
        keys[o] = key;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, d));
    if (threadIdx.x == 0 && m) atomicMax(disp_bits, m);
}

__global__ void __launch_bounds__(256) k_scatter_ordered(Geom g, const float* __restrict__ src, const float* __restrict__ u,
                                                        const float* __restrict__ v, const unsigned* __restrict__ keys,
                                                        float* __restrict__ out, float dt0, const unsigned* __restrict__ disp_bits) {
    const int tj = 1 + blockIdx.x * 32 + threadIdx.x, ti = 1 + blockIdx.y * 8 + threadIdx.y;
    if (ti > g.rows - 2 || tj > g.cols - 2) return;
    const int R = sc::reach(*disp_bits, max(g.rows, g.cols));
    const int ilo = max(1, ti - R), ihi = min(g.rows - 2, ti + R);
    const int jlo = max(1, tj - R), jhi = min(g.cols - 2, tj + R);
    const unsigned P = (unsigned)g.pitch, T = (unsigned)ti * P + (unsigned)tj;
    float acc = 0.f;
    for (int i = ilo; i <= ihi; ++i) {
        const size_t row = (size_t)i * g.pitch;
        for (int j = jlo; j <= jhi; ++j) {
            const unsigned d = T - __ldg(keys + row + j);
            if (!sc::is_hit(d, P)) continue;
            acc = __fadd_rn(acc, sc::share(g.rows, g.cols, i, j, __ldg(u + row + j), __ldg(v + row + j), dt0, d, P, __ldg(src + row + j)));
        }
    }
    out[(size_t)ti * g.pitch + tj] = acc;
}

void launch_scatter_ordered(const Geom& g, const float* src, const float* u, const float* v, float* out, unsigned* keys,
                            float dt0, unsigned* disp_bits, cudaStream_t st) {
    cudaMemsetAsync(disp_bits, 0, sizeof(unsigned), st);
    const dim3 bl(32, 8);
    k_scatter_keys<<<dim3((g.cols + 31) / 32, (g.rows + 7) / 8), bl, 0, st>>>(g, u, v, dt0, keys, disp_bits);
    k_scatter_ordered<<<dim3((g.cols - 2 + 31) / 32, (g.rows - 2 + 7) / 8), bl, 0, st>>>(g, src, u, v, keys, out, dt0, disp_bits);
}

}  // namespace f2d
