// f2d_jacobi_stream.cu -- temporally blocked Jacobi relaxation for sm_100a.
//
// Replaces the reference's relaxation loops (diffuse: src/fluid_solver_gpu.cu:301-311 with
// diffuse_iteration_kernel :69-85; pressure: :379-391 with p_iteration_kernel :179-191), which
// per sweep do a full-field D2D copy, one 2-D kernel, one 1-D boundary kernel and two device
// syncs, i.e. >= 20 B of HBM traffic per cell per sweep.
//
// Design ("row streaming with a register pipeline"):
//   * The field is cut into column strips of 128 floats (one float4 per lane of a warp) and row
//     chunks.  ONE WARP owns one (strip, chunk) and marches down its rows.  No block-level
//     synchronisation exists anywhere: warps are fully independent.
//   * T sweeps are fused per pass.  For every time level s < T a lane keeps a sliding window of
//     three rows (its four columns) in REGISTERS; when input row r arrives, level 1 can produce
//     row r-1, level 2 row r-2, ... level T row r-T, which is stored.  North/south neighbours are
//     therefore register reads, west/east neighbours are two warp shuffles per row, and each
//     input row is read from HBM once and each output row written once per T sweeps
//     (12/T bytes per cell-sweep instead of 12).
//   * Redundant work exists only in the HALO columns left/right of a strip (HALO >= T) and the
//     T warm-up rows above/below a chunk; validity shrinks by one cell per level from every
//     non-domain edge, exactly covered by the halos.
//   * Input rows are staged through a per-warp shared-memory ring with 16-byte asynchronous
//     copies (cp.async.cg -> LDGSTS, L1-bypassing), PFD rows ahead, so HBM latency is hidden
//     without spending registers; each lane only ever reads back the bytes it copied itself,
//     so cp.async.wait_group is the only synchronisation needed.  Ring slots are addressed on
//     32-bit shared addresses as ((row << 9) & mask) | aligned_base (two integer instructions).
//   * The boundary pass (set_boundary_*, gpu.cu:11-54) is fused: domain edge columns are fixed
//     inside the lane that holds them (columns 0/1 and N-2/N-1 share a float4 because cols%4==0),
//     edge rows are produced by the edge rule when the adjacent interior row of the same level
//     is produced; corners are carried through unchanged from the input (the reference never
//     writes them, gpu.cu:15-23).
//   * Arithmetic is spelled with intrinsics (f2d_common.cuh) so every level is bit-identical to
//     one sweep of the naive kernel: T fused sweeps == T single sweeps, bitwise.
//   * The relaxation itself is issued as PACKED fp32x2 instructions (add/mul/fma.rn.f32x2 -> FADD2 / FMUL2 / FFMA2,
//     new on sm_100): a lane keeps its four columns (x, y, z, w) as the two register pairs A = (x, z) and B = (y, w).
//     With that pairing every operand of a packed operation is an aligned pair already: north / south / rhs are the
//     same pair of another row, the east neighbours of A are B, the west neighbours of B are A; only the west
//     neighbours of A = (left lane's w, y) and the east neighbours of B = (z, right lane's x) are assembled, one
//     shuffle + one move each.  Each half of a packed operation is the IEEE operation of the scalar kernel, so the
//     bits do not change; the instruction count per row and level drops from 28 to 17 (pressure) and from 39 to 23
//     (diffuse).  FADD2 occupies the FMA pipe for two cycles (profiles/ubench_fp32x2_r02.jsonl): what is won are
//     issue slots, which is what bound the scalar kernel (profiles/ncu_jacobi_*_T8_r01_final.md).  Right-hand-side
//     rows are re-written in the smem ring in (x, z, y, w) order once when they land, so that every later read is
//     one 16-byte load straight into two pairs.
//   * The edge-column fix (columns 0 / N-1 = +/- their neighbour) is compiled only into the variant run by the two
//     strips that hold a domain edge column (template parameter EDGE, warp-uniform choice outside the row loop).
//   * The row loop is unrolled by RS (a multiple of 3) so that all window/ring register indices
//     are compile-time constants; a row step is one basic block; blocks of RS rows in which no
//     level meets a GLOBAL edge row take a FAST path without range or edge-row checks.
//   * Two fused first passes (template parameter PIN_ZERO): == 2, the first pressure pass computes
//     the divergence from u and v on the fly (p0 == 0) and writes it out for the later passes;
//     == 3, the first diffuse pass forms x0 = FMA(dt, source, field) (add_sources), relaxes from it
//     and writes it out as the right-hand side of the later passes.
//   * A host planner (launch_one) sizes the chunks from a cost model so that one launch is exactly
//     one wave of resident warps that finish together.
#include <algorithm>

#include "f2d_kernels.cuh"

// A/B switches (tools/build_variants.sh builds one library per combination)
#ifndef F2D_SHFL_AHEAD
#define F2D_SHFL_AHEAD 0  // 1: west/east shuffles issued at the end of the previous row step (costs 2T live
                          // registers; slower since the row step became one basic block, profiles/ab_r01_run7_*.log)
#endif

namespace f2d {

namespace {

constexpr int kLanes = 32;
constexpr int kStripFloats = 128;  // one float4 per lane
constexpr int kPFD = 6;            // async prefetch distance in rows
constexpr int kRingP = 8;          // ring slots for the iterate rows (power of two, >= PFD + 2)

__host__ __device__ constexpr int halo_of(int T) { return T <= 4 ? 4 : ((T + 3) / 4) * 4; }
__host__ __device__ constexpr int m3(int x) { return ((x % 3) + 3) % 3; }
__host__ __device__ constexpr int rs_of(int T) { return 3 * ((T + 1 + 2) / 3); }  // rhs register ring
__host__ __device__ constexpr int mrs(int x, int RS) { return ((x % RS) + RS) % RS; }
// rhs smem ring slots (power of two): register mode only needs the landing zone,
// smem mode keeps rows r+PFD .. r-T
__host__ __device__ constexpr int ring_r_of(int T, bool rhs_regs) {
    return rhs_regs ? kRingP : ((kPFD + T + 2) <= 8 ? 8 : 16);
}

// Work decomposition.  Warps fall into two classes with different cost per row: class 0 = interior
// strips, class 1 = the strips that hold a left/right domain edge column (extra edge fix per level).
// Each class has its own chunk height so that all warps of a launch finish together; the first and
// last chunk of a class are `edge_trim` rows shorter to pay for their checked edge-row steps.
struct StreamPlan {
    int strips, bw;
    int warps_int;      // warps of class 0: (strips - n_edge_strips) * chunks[0]
    int n_edge_strips;  // 0, 1 or 2
    int chunks[2], chunk_rows[2], edge_trim[2];
};

__device__ __forceinline__ void cp_async16_s(unsigned smem_addr, const void* gmem_src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_addr), "l"(gmem_src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ float4 lds128(unsigned smem_addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts128(unsigned smem_addr, const float4& v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"r"(smem_addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void st_global_f4(float* p, const float4& v) {
    asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// ---- packed fp32x2 values: two floats in one aligned 64-bit register pair (lo | hi << 32)
typedef unsigned long long f2;
__device__ __forceinline__ f2 pk(float lo, float hi) {
    f2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float lo_of(f2 v) {
    float a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
    return a;
}
__device__ __forceinline__ float hi_of(f2 v) {
    float a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
    return b;
}
// each half is the IEEE round-to-nearest operation of the scalar kernel (no flush-to-zero, no contraction)
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
    f2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
    f2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
    f2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// one row of a lane: its four columns (x, y, z, w) held as A = (x, z), B = (y, w)
struct P {
    f2 A, B;
};
__device__ __forceinline__ P perm(const float4& v) {
    P p;
    p.A = pk(v.x, v.z);
    p.B = pk(v.y, v.w);
    return p;
}
__device__ __forceinline__ float4 unperm(const P& p) { return make_float4(lo_of(p.A), lo_of(p.B), hi_of(p.A), hi_of(p.B)); }
__device__ __forceinline__ P zero_p() {
    P p;
    p.A = p.B = 0ull;
    return p;
}
// a ring slot that holds a row in (x, z, y, w) order: 16 bytes <-> two pairs
__device__ __forceinline__ P lds_p(unsigned smem_addr) {
    P p;
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];\n" : "=l"(p.A), "=l"(p.B) : "r"(smem_addr) : "memory");
    return p;
}
__device__ __forceinline__ void sts_p(unsigned smem_addr, const P& p) {
    asm volatile("st.shared.v2.b64 [%0], {%1, %2};\n" ::"r"(smem_addr), "l"(p.A), "l"(p.B) : "memory");
}

// the coefficients of one relaxation, broadcast into pairs
struct Coef2 {
    f2 a, rc, nch, ncl;  // diffuse: a, RN32(1/c), -ch, -cl (c = ch + cl)
    f2 q;                // pressure: 0.25
    DiffuseCoef s;       // scalar copy (fp64-divide mode)
};
__device__ __forceinline__ Coef2 make_coef2(const DiffuseCoef& k) {
    Coef2 c;
    c.a = pk(k.a, k.a);
    c.rc = pk(k.rc, k.rc);
    c.nch = pk(-k.ch, -k.ch);
    c.ncl = pk(-k.cl, -k.cl);
    c.q = pk(0.25f, 0.25f);
    c.s = k;
    return c;
}

// diffuse_iteration_kernel (src/fluid_solver_gpu.cu:81-82) for two cells, fp32-corrected divide; the same operations as
// diffuse_update<F2D_DIV_F32_CORR> (f2d_common.cuh): FMA(-q0, ch, num) == FMA(q0, -ch, num) exactly
__device__ __forceinline__ f2 diffuse2(f2 w, f2 e, f2 n, f2 sth, f2 x0, const Coef2& k) {
    const f2 sum = add2(add2(add2(w, e), n), sth);
    const f2 num = fma2(k.a, sum, x0);
    const f2 q0 = mul2(num, k.rc);
    f2 r = fma2(q0, k.nch, num);
    r = fma2(q0, k.ncl, r);
    return fma2(r, k.rc, q0);
}
// p_iteration_kernel (src/fluid_solver_gpu.cu:187-188) for two cells
__device__ __forceinline__ f2 pressure2(f2 dv, f2 e, f2 w, f2 sth, f2 n, f2 quarter) {
    return mul2(add2(add2(add2(add2(dv, e), w), sth), n), quarter);
}

// level s+1 row from the level-s rows a (north), b (centre), c (south); l / rt = the cells left of x / right of w
template <bool DIFFUSE, int DIVMODE>
__device__ __forceinline__ P relax_row(const P& a, const P& b, const P& c, float l, float rt, const P& rhs, const Coef2& k) {
    P o;
    if (DIFFUSE && DIVMODE == F2D_DIV_F64) {  // the reference's fp64 divide: scalar
        const float4 av = unperm(a), bv = unperm(b), cv = unperm(c), rv = unperm(rhs);
        float4 ov;
        ov.x = diffuse_update<F2D_DIV_F64>(l, bv.y, av.x, cv.x, rv.x, k.s);
        ov.y = diffuse_update<F2D_DIV_F64>(bv.x, bv.z, av.y, cv.y, rv.y, k.s);
        ov.z = diffuse_update<F2D_DIV_F64>(bv.y, bv.w, av.z, cv.z, rv.z, k.s);
        ov.w = diffuse_update<F2D_DIV_F64>(bv.z, rt, av.w, cv.w, rv.w, k.s);
        return perm(ov);
    }
    const f2 WL = pk(l, lo_of(b.B));   // west of (x, z) = (left lane's w, y)
    const f2 ER = pk(hi_of(b.A), rt);  // east of (y, w) = (z, right lane's x)
    if (DIFFUSE) {
        o.A = diffuse2(WL, b.B, a.A, c.A, rhs.A, k);
        o.B = diffuse2(b.A, ER, a.B, c.B, rhs.B, k);
    } else {
        o.A = pressure2(rhs.A, b.B, WL, c.A, a.A, k.q);
        o.B = pressure2(rhs.B, ER, b.A, c.B, a.B, k.q);
    }
    return o;
}

// domain edge columns of an interior row: col 0 = +/- col 1, col N-1 = +/- col N-2 (gpu.cu:16-17, 31-32)
__device__ __forceinline__ P fix_edge_cols(const P& v, bool has_left, bool has_right, bool neg) {
    float4 f = unperm(v);
    f.x = has_left ? apply_sign(f.y, neg) : f.x;
    f.w = has_right ? apply_sign(f.z, neg) : f.w;
    return perm(f);
}

// edge row from the adjacent interior row of the same level; corner cells keep `keep`
__device__ __forceinline__ P edge_row(const P& inner_p, const P& keep_p, bool neg, bool has_left, bool has_right) {
    const float4 inner = unperm(inner_p), keep = unperm(keep_p);
    float4 o;
    o.x = apply_sign(inner.x, neg);
    o.y = apply_sign(inner.y, neg);
    o.z = apply_sign(inner.z, neg);
    o.w = apply_sign(inner.w, neg);
    if (has_left) o.x = keep.x;
    if (has_right) o.w = keep.w;
    return perm(o);
}

// per-warp constants of one (strip, chunk)
struct Ctx {
    const float* prev;  // lane-adjusted: + column of this lane
    const float* rhs;
    float* next;
    unsigned sp, sr;  // the two async-copy rings as 32-bit shared addresses (lane offset included), each
                      // aligned to its own size so that a slot is ((row << 9) & mask) | base
    unsigned sd;      // fused divergence: ring of computed divergence rows (the relaxation's rhs)
    float* aux;       // fused divergence: the divergence field written for the later passes
    float mhalf_h;    // fused divergence: -0.5f * h; fused add_sources: dt
    int row_top, row_bot;  // local index of the global top / bottom edge row (or out of range)
    Coef2 coef;
    int pitch, rs, re, y0, y1;
    int cp_bytes;
    bool top_dom, bot_dom, own_x, has_left, has_right, edge_warp, neg_c, neg_r;
};

// RS consecutive row steps starting at relative row rb (a multiple of RS, so rb % 3 == 0 and the
// register slots of every row are compile-time constants).
//
// FAST blocks run every level unconditionally.  During the warm-up / drain of a chunk some levels
// then work on rows outside the chunk's input range (garbage, but finite): those results only ever
// feed cells outside the dependency cone of the rows this warp stores, and the store itself is
// predicated on the owned row range.  Only blocks in which a level meets a GLOBAL top or bottom edge
// row (edge rule, corner carry) take the checked path (FAST == false).
template <int T, bool DIFFUSE, int DIVMODE, int PIN_ZERO, bool RHS_REGS, bool FAST, bool EDGE, int RS, int RINGR, int NRH>
__device__ __forceinline__ void run_block(const Ctx& cx, int rb, int nsteps, P (&W)[T][3], P (&RH)[NRH],
                                          P& out_prev, float (&wl)[T], float (&er)[T], float4 (&UV)[2][3]) {
    // PIN_ZERO == 2: the first pressure pass with the divergence fused in (gpu.cu:164-177 + :376): the two
    // async rings carry u and v rows instead of iterate and rhs; the rhs row r-1 is computed on the fly
    constexpr bool FUSE = (PIN_ZERO == 2);
    // PIN_ZERO == 3: the first diffuse pass with add_sources fused in (gpu.cu:56-67 folded into :290-312): the
    // rings carry the field and its source; x0 = FMA(dt, s, f) is formed per row, used as iterate AND rhs,
    // and written out as the rhs of the later passes
    constexpr bool FSRC = (PIN_ZERO == 3);
    constexpr unsigned MASKR = (FUSE || FSRC) ? ((kRingP - 1) << 9) : ((RINGR - 1) << 9);  // landing zone only
#pragma unroll
    for (int k = 0; k < RS; ++k) {
        const int rr = rb + k;
        if (rr >= nsteps) break;
        const int r = cx.rs + rr;

        // 1. keep PFD rows in flight
        {
            const int rl = r + kPFD;
            if (rl <= cx.re) {
                const size_t off = (size_t)rl * cx.pitch;
                if (PIN_ZERO != 1) cp_async16_s((((unsigned)rl << 9) & ((kRingP - 1) << 9)) | cx.sp, cx.prev + off, cx.cp_bytes);
                cp_async16_s((((unsigned)rl << 9) & MASKR) | cx.sr, cx.rhs + off, cx.cp_bytes);
            }
            cp_async_commit();
        }
        // 2. row r has landed (each lane reads back only the 16 bytes it copied itself)
        cp_async_wait<kPFD>();
        if (FAST || r <= cx.re) {  // FAST: past the last input row this re-reads a stale ring slot (harmless)
            if (FSRC) {
                const float4 f4 = lds128((((unsigned)r << 9) & ((kRingP - 1) << 9)) | cx.sp);
                const float4 s4 = lds128((((unsigned)r << 9) & MASKR) | cx.sr);
                const bool row_in = (r != cx.row_top) && (r != cx.row_bot);  // global interior row
                float4 x0;
                x0.x = (row_in && !cx.has_left) ? __fmaf_rn(cx.mhalf_h, s4.x, f4.x) : f4.x;  // mhalf_h carries dt here
                x0.y = row_in ? __fmaf_rn(cx.mhalf_h, s4.y, f4.y) : f4.y;
                x0.z = row_in ? __fmaf_rn(cx.mhalf_h, s4.z, f4.z) : f4.z;
                x0.w = (row_in && !cx.has_right) ? __fmaf_rn(cx.mhalf_h, s4.w, f4.w) : f4.w;
                const P x0p = perm(x0);
                W[0][m3(k)] = x0p;
                if (RHS_REGS)
                    RH[mrs(k, RS)] = x0p;
                else
                    sts_p((((unsigned)r << 9) & ((RINGR - 1) << 9)) | cx.sd, x0p);
                if (cx.own_x && r >= cx.y0 && r < cx.y1 && r <= cx.re) st_global_f4(cx.aux + (size_t)r * cx.pitch, x0);
            } else if (!PIN_ZERO)
                W[0][m3(k)] = perm(lds128((((unsigned)r << 9) & ((kRingP - 1) << 9)) | cx.sp));
            else
                W[0][m3(k)] = zero_p();
            if (FSRC) {
            } else if (FUSE) {
                UV[0][m3(k)] = lds128((((unsigned)r << 9) & ((kRingP - 1) << 9)) | cx.sp);  // u row r
                UV[1][m3(k)] = lds128((((unsigned)r << 9) & MASKR) | cx.sr);                // v row r
            } else if (RHS_REGS) {
                RH[mrs(k, RS)] = perm(lds128((((unsigned)r << 9) & ((RINGR - 1) << 9)) | cx.sr));
            } else {
                // the rhs row just landed in (x, y, z, w) order: re-write it as (x, z, y, w) so that each of its T
                // later reads is one 16-byte load into two aligned pairs (every lane touches only its own 16 bytes)
                const unsigned slot = (((unsigned)r << 9) & ((RINGR - 1) << 9)) | cx.sr;
                sts_p(slot, perm(lds128(slot)));
            }
        }
        if (FUSE) {
            // divergence of row r-1 from u[r-1] (with its west/east neighbours) and v[r], v[r-2]
            const float4 uc = UV[0][m3(k - 1)], vs = UV[1][m3(k)], vn = UV[1][m3(k - 2)];
            const float ul = __shfl_up_sync(0xffffffffu, uc.w, 1);
            const float ur = __shfl_down_sync(0xffffffffu, uc.x, 1);
            float4 dv;
            dv.x = divergence_update(uc.y, ul, vs.x, vn.x, cx.mhalf_h);
            dv.y = divergence_update(uc.z, uc.x, vs.y, vn.y, cx.mhalf_h);
            dv.z = divergence_update(uc.w, uc.y, vs.z, vn.z, cx.mhalf_h);
            dv.w = divergence_update(ur, uc.z, vs.w, vn.w, cx.mhalf_h);
            // set_boundary_continuous on the divergence (gpu.cu:376): edge columns copy their neighbour
            dv.x = cx.has_left ? dv.y : dv.x;
            dv.w = cx.has_right ? dv.z : dv.w;
            const int qd = r - 1;
            if (RHS_REGS)
                RH[mrs(k - 1, RS)] = perm(dv);
            else
                sts_p((((unsigned)qd << 9) & ((RINGR - 1) << 9)) | cx.sd, perm(dv));
            if (cx.own_x && qd >= cx.rs + 1 && qd <= cx.re - 1) {
                if (qd >= cx.y0 && qd < cx.y1) st_global_f4(cx.aux + (size_t)qd * cx.pitch, dv);
                // the edge rows of the stored field copy the adjacent interior row, corners are the memset zeros
                float4 e = dv;
                e.x = cx.has_left ? 0.f : e.x;
                e.w = cx.has_right ? 0.f : e.w;
                if (qd == 1 && cx.top_dom && cx.y0 == 0) st_global_f4(cx.aux, e);
                if (qd == cx.re - 1 && cx.bot_dom) st_global_f4(cx.aux + (size_t)cx.re * cx.pitch, e);
            }
        }
#if !F2D_SHFL_AHEAD
        // west/east neighbours of the centre rows of all levels (rows produced in the previous step)
#pragma unroll
        for (int s = 0; s < T; ++s) {
            const P bb = W[s][m3(k - s - 1)];
            wl[s] = __shfl_up_sync(0xffffffffu, hi_of(bb.B), 1);    // the left lane's w
            er[s] = __shfl_down_sync(0xffffffffu, lo_of(bb.A), 1);  // the right lane's x
        }
#endif

        // 3. level s+1 produces row q = r - s - 1 from level s rows q-1, q, q+1; wl[s], er[s] are the
        //    west/east neighbours of the centre row (a row produced in the previous step).
        //    active levels: rs+1 <= q <= re-1  <=>  s_lo <= s <= s_hi
        const int s_lo = r - cx.re, s_hi = r - cx.rs - 2;
        const int s_top = cx.top_dom ? r - 2 : -1;           // level whose q == 1 (global top edge above it)
        const int s_bot = cx.bot_dom ? r - cx.re - 1 : -1;   // level whose q == re == global bottom edge row
#pragma unroll
        for (int s = 0; s < T; ++s) {
            const int q = r - s - 1;
            const int sa = m3(k - s - 2), sm = m3(k - s - 1), sc = m3(k - s);
            const int sn = (s + 1 < T) ? s + 1 : 0;  // keeps the dead branch's index in range
            if (FAST || (s >= s_lo && s <= s_hi)) {
                const P a = W[s][sa], b = W[s][sm], c = W[s][sc];
                const float l = wl[s], rt = er[s];
                P rhs;
                if (RHS_REGS)
                    rhs = RH[mrs(k - s - 1, RS)];
                else
                    rhs = lds_p((((unsigned)q << 9) & ((RINGR - 1) << 9)) | ((FUSE || FSRC) ? cx.sd : cx.sr));
                P nw = relax_row<DIFFUSE, DIVMODE>(a, b, c, l, rt, rhs, cx.coef);
                // only the strips that hold a domain edge column run this variant; branch-free selects, so a whole
                // row step stays one basic block
                if (EDGE) nw = fix_edge_cols(nw, cx.has_left, cx.has_right, cx.neg_c);
                if (s + 1 < T) {
                    W[sn][sm] = nw;
                } else {
                    out_prev = nw;
                    if (cx.own_x && q >= cx.y0 && q < cx.y1) st_global_f4(cx.next + (size_t)q * cx.pitch, unperm(nw));
                }
                if (!FAST && s == s_top) {  // global top edge row of the same level (corners kept)
                    const P e = edge_row(nw, a, cx.neg_r, cx.has_left, cx.has_right);
                    if (s + 1 < T)
                        W[sn][sa] = e;
                    else if (cx.own_x && cx.y0 == 0)
                        st_global_f4(cx.next, unperm(e));
                }
            } else if (!FAST && s == s_bot && q >= cx.rs + 1) {  // global bottom edge row
                const P inner = (s + 1 < T) ? W[sn][sa] : out_prev;
                const P e = edge_row(inner, W[s][sm], cx.neg_r, cx.has_left, cx.has_right);
                if (s + 1 < T)
                    W[sn][sm] = e;
                else if (cx.own_x && q >= cx.y0 && q < cx.y1)
                    st_global_f4(cx.next + (size_t)q * cx.pitch, unperm(e));
            }
        }
        // 4. the centre row of level s in the NEXT step is row r - s (slot m3(k - s)); it is final now,
        //    so its west/east shuffles are issued here and complete while the next row is fetched
#if F2D_SHFL_AHEAD
#pragma unroll
        for (int s = 0; s < T; ++s) {
            const P b = W[s][m3(k - s)];
            wl[s] = __shfl_up_sync(0xffffffffu, hi_of(b.B), 1);
            er[s] = __shfl_down_sync(0xffffffffu, lo_of(b.A), 1);
        }
#endif
    }
}

// MINB = resident CTAs (of 128 threads) per SM the register allocator must allow
template <int T, bool DIFFUSE, int DIVMODE, int PIN_ZERO, bool RHS_REGS, int MINB>
__global__ void __launch_bounds__(128, MINB) k_jacobi_stream(Geom g, RelaxBatch batch, StreamPlan plan) {
    constexpr int HALO = halo_of(T);
    constexpr int RS = RHS_REGS ? rs_of(T) : 3;  // unroll factor of the row loop
    constexpr int RINGR = ring_r_of(T, RHS_REGS);
    constexpr int NRH = RHS_REGS ? RS : 1;
    extern __shared__ float4 smem[];

    const int lane = threadIdx.x & 31;
    const int warp_in_cta = threadIdx.x >> 5;
    const int warps_per_cta = blockDim.x >> 5;
    const int gw = blockIdx.x * warps_per_cta + warp_in_cta;
    int cls, strip, chunk;
    if (gw < plan.warps_int) {  // interior strips: 1 .. strips-2 (or 0 .. when there is no edge strip)
        cls = 0;
        const int ns = plan.strips - plan.n_edge_strips;
        strip = (plan.n_edge_strips ? 1 : 0) + gw % ns;
        chunk = gw / ns;
    } else {  // the strips holding column 0 / column cols-1
        cls = 1;
        const int e = gw - plan.warps_int;
        if (plan.n_edge_strips == 0) return;
        strip = (e % plan.n_edge_strips == 0) ? 0 : plan.strips - 1;
        chunk = e / plan.n_edge_strips;
    }
    const int n_chunks = plan.chunks[cls], chunk_rows = plan.chunk_rows[cls], edge_trim = plan.edge_trim[cls];
    if (chunk >= n_chunks) return;  // whole warp leaves; there is no block-wide barrier

    const RelaxField& fld = batch.f[blockIdx.y];
    Ctx cx;
    cx.coef = make_coef2(fld.coef);
    cx.neg_c = (fld.kind == F2D_BND_OPPOSITE_HORIZONTAL);
    cx.neg_r = (fld.kind == F2D_BND_OPPOSITE_VERTICAL);
    cx.pitch = g.pitch;

    // ---- columns of this lane
    const int jb = strip * plan.bw - HALO + 4 * lane;
    const bool in_dom = (jb >= 0) && (jb + 3 < g.cols);
    cx.has_left = (jb == 0);
    cx.has_right = (jb + 3 == g.cols - 1);
    cx.edge_warp = __any_sync(0xffffffffu, cx.has_left || cx.has_right) != 0;
    cx.own_x = in_dom && (lane >= HALO / 4) && (lane < kLanes - HALO / 4);
    cx.cp_bytes = in_dom ? 16 : 0;
    const int jsafe = in_dom ? jb : 0;
    cx.prev = (PIN_ZERO == 1) ? nullptr : fld.prev + jsafe;  // PIN_ZERO == 2: u
    cx.aux = (PIN_ZERO >= 2) ? fld.aux + jsafe : nullptr;
    cx.mhalf_h = (PIN_ZERO == 3) ? batch.dt : fld.coef.a;  // fused divergence: the launcher passes -0.5f*h in coef.a
    cx.row_top = (g.grow0 == 0) ? 0 : -1;
    cx.row_bot = g.grows - 1 - g.grow0;
    cx.rhs = fld.rhs + jsafe;
    cx.next = fld.next + jsafe;
    {
        // [ring_r of warp 0 .. wpc-1][ring_p of warp 0 .. wpc-1], the block aligned to the rhs ring size so
        // that "(row << 9) & mask | base" addresses a slot with two integer instructions
        constexpr unsigned RB = RINGR * kLanes * 16u, PB = kRingP * kLanes * 16u;
        const unsigned s0 = ((unsigned)__cvta_generic_to_shared(smem) + RB - 1u) & ~(RB - 1u);
        if (PIN_ZERO >= 2) {
            // fused divergence / add_sources: [computed rhs ring (RB) x wpc][landing ring (PB) x wpc][landing ring (PB) x wpc]
            cx.sd = s0 + (unsigned)warp_in_cta * RB + (unsigned)lane * 16u;
            cx.sr = s0 + (unsigned)warps_per_cta * RB + (unsigned)warp_in_cta * PB + (unsigned)lane * 16u;
            cx.sp = s0 + (unsigned)warps_per_cta * (RB + PB) + (unsigned)warp_in_cta * PB + (unsigned)lane * 16u;
        } else {
            cx.sd = 0;
            cx.sr = s0 + (unsigned)warp_in_cta * RB + (unsigned)lane * 16u;
            cx.sp = s0 + (unsigned)warps_per_cta * RB + (unsigned)warp_in_cta * PB + (unsigned)lane * 16u;
        }
    }

    // ---- rows of this warp (local row indices)
    // the first / last chunk are `edge_trim` rows shorter: their edge-rule steps cost more
    cx.y0 = (chunk == 0) ? 0 : chunk * chunk_rows - edge_trim;
    cx.y1 = (chunk == n_chunks - 1) ? g.rows : (chunk + 1) * chunk_rows - edge_trim;
    // T warm-up rows above the first owned row; a chunk that owns only the global bottom edge row
    // must warm up for row rows-2, which the edge rule copies from
    cx.rs = max(0, min(cx.y0, g.rows - 2) - T);
    cx.re = min(g.rows - 1, cx.y1 - 1 + T);
    // pin the per-warp constants in registers: without this the compiler re-derives them every row
    // from the (dynamically indexed) kernel-parameter bank
    asm volatile("" : "+l"(cx.prev), "+l"(cx.rhs), "+l"(cx.next));
    asm volatile("" : "+r"(cx.pitch), "+r"(cx.rs), "+r"(cx.re), "+r"(cx.y0), "+r"(cx.y1), "+r"(cx.cp_bytes));
    cx.top_dom = (cx.rs == 0) && (g.grow0 == 0);
    cx.bot_dom = (cx.re == g.rows - 1) && (g.grow0 + g.rows == g.grows);
    const int nsteps = (cx.y1 - 1 + T) - cx.rs + 1;
    // absolute input rows r during which some level meets a global edge row:
    //   top:    q == 1 at level s+1  <=>  r = s + 2,      s in [0, T)
    //   bottom: q == re at level s+1 <=>  r = re + s + 1, s in [0, T)
    const int top_lo = cx.top_dom ? 2 : 1 << 30, top_hi = cx.top_dom ? T + 1 : -1;
    const int bot_lo = cx.bot_dom ? cx.re + 1 : 1 << 30, bot_hi = cx.bot_dom ? cx.re + T : -1;

    P W[T][3];
    P RH[NRH];
    P out_prev = zero_p();
    float wl[T], er[T];
    float4 UV[2][3];
#pragma unroll
    for (int m = 0; m < 3; ++m) UV[0][m] = UV[1][m] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int s = 0; s < T; ++s) {
        wl[s] = er[s] = 0.f;
#pragma unroll
        for (int m = 0; m < 3; ++m) W[s][m] = zero_p();
    }
#pragma unroll
    for (int m = 0; m < NRH; ++m) RH[m] = zero_p();

    // ---- prologue: rows rs .. rs+PFD-1 in flight
#pragma unroll
    for (int p = 0; p < kPFD; ++p) {
        const int rl = cx.rs + p;
        if (rl <= cx.re) {
            const size_t off = (size_t)rl * cx.pitch;
            if (PIN_ZERO != 1) cp_async16_s((((unsigned)rl << 9) & ((kRingP - 1) << 9)) | cx.sp, cx.prev + off, cx.cp_bytes);
            cp_async16_s((((unsigned)rl << 9) & ((PIN_ZERO >= 2 ? kRingP - 1 : RINGR - 1) << 9)) | cx.sr, cx.rhs + off, cx.cp_bytes);
        }
        cp_async_commit();
    }

    // the edge-column variant is chosen once per warp (warp-uniform), outside the row loop
    if (!cx.edge_warp) {
        for (int rb = 0; rb < nsteps; rb += RS) {
            const int r_first = cx.rs + rb, r_last = r_first + RS - 1;
            const bool edge_block = (r_first <= top_hi && r_last >= top_lo) || (r_first <= bot_hi && r_last >= bot_lo);
            if (!edge_block)
                run_block<T, DIFFUSE, DIVMODE, PIN_ZERO, RHS_REGS, true, false, RS, RINGR, NRH>(cx, rb, nsteps, W, RH, out_prev, wl, er, UV);
            else
                run_block<T, DIFFUSE, DIVMODE, PIN_ZERO, RHS_REGS, false, false, RS, RINGR, NRH>(cx, rb, nsteps, W, RH, out_prev, wl, er, UV);
        }
    } else {
        for (int rb = 0; rb < nsteps; rb += RS) {
            const int r_first = cx.rs + rb, r_last = r_first + RS - 1;
            const bool edge_block = (r_first <= top_hi && r_last >= top_lo) || (r_first <= bot_hi && r_last >= bot_lo);
            if (!edge_block)
                run_block<T, DIFFUSE, DIVMODE, PIN_ZERO, RHS_REGS, true, true, RS, RINGR, NRH>(cx, rb, nsteps, W, RH, out_prev, wl, er, UV);
            else
                run_block<T, DIFFUSE, DIVMODE, PIN_ZERO, RHS_REGS, false, true, RS, RINGR, NRH>(cx, rb, nsteps, W, RH, out_prev, wl, er, UV);
        }
    }
    cp_async_wait<0>();
}

// Host-side planner.  One wave, every resident warp slot busy, all warps finishing together:
//   cost(warp) = (rows of its chunk + 2T warm-up rows) * c_class  [+ checked edge-row steps]
// with c_class the instruction count of one row step (interior strip : edge strip ~ 1 : kEdgeStripCost,
// read off the SASS).  The largest chunk heights whose warp count fits the resident slots are found
// by bisection on the common cost.  First/last chunks are trimmed by the cost of their checked steps.
constexpr double kEdgeStripCost = 1.35;  // default cost of an edge strip's row step relative to an interior strip's (SASS count;
                                         // only the edge strips carry the edge-column fix); F2D_STREAM_EDGE_COST_PCT overrides
constexpr double kCheckedStepCost = 3.4; // checked (global edge row) step relative to a fast step

struct ClassPlan {
    int chunks, chunk_rows, trim;
};

inline ClassPlan plan_class(int rows, int T, int RS, double cost_budget, double c_class, bool both_edges, int min_mult) {
    ClassPlan cp;
    int ch = (int)(cost_budget / c_class) - 2 * T;  // rows per chunk at this budget
    ch = std::max(ch, min_mult * T);  // bound the warm-up redundancy on small grids
    ch = std::min(ch, rows);
    int trim = both_edges ? (int)((T + RS / 2) * (kCheckedStepCost - 1.0)) : 0;
    if (ch < 2 * trim + 2 || ch >= rows) trim = 0;
    int chunks = (rows + 2 * trim + ch - 1) / ch;
    while (chunks > 1 && (chunks - 1) * ch - trim >= rows) --chunks;
    if (chunks < 3) {
        trim = 0;
        chunks = (rows + ch - 1) / ch;
    }
    cp.chunks = chunks;
    cp.chunk_rows = ch;
    cp.trim = trim;
    return cp;
}

template <int T, bool DIFFUSE, int DIVMODE, int PIN_ZERO, bool RHS_REGS, int MINB>
cudaError_t launch_one(const Geom& g, const RelaxBatch& b, const StreamTuning& tune, int sm_count, cudaStream_t st) {
    auto kern = k_jacobi_stream<T, DIFFUSE, DIVMODE, PIN_ZERO, RHS_REGS, MINB>;
    constexpr int RS = RHS_REGS ? rs_of(T) : 3;
    int wpc = tune.warps_per_cta > 0 ? tune.warps_per_cta : 4;
    wpc = std::min(wpc, 4);  // __launch_bounds__(128, ...)
    const size_t smem = (size_t)wpc * (kRingP + ring_r_of(T, RHS_REGS) + (PIN_ZERO >= 2 ? kRingP : 0)) * kLanes * sizeof(float4) +
                        (size_t)ring_r_of(T, RHS_REGS) * kLanes * sizeof(float4);  // alignment slack
    // per device (function attributes live in the device's context) and per CTA size; a failed opt-in to more than
    // 48 KB of dynamic shared memory or a failed occupancy query is reported here, not at some later launch
    static int occ_cache_dev[16][9] = {};
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return err;
    int* occ_cache = occ_cache_dev[dev & 15];
    if (occ_cache[wpc] == 0) {
        if (smem > 48 * 1024) {
            err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (err != cudaSuccess) return err;
        }
        int occ = 0;
        err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, wpc * 32, smem);
        if (err != cudaSuccess) return err;
        if (occ < 1) return cudaErrorLaunchOutOfResources;
        occ_cache[wpc] = occ;
    }
    StreamPlan plan;
    plan.bw = kStripFloats - 2 * halo_of(T);
    plan.strips = (g.cols + plan.bw - 1) / plan.bw;
    plan.n_edge_strips = std::min(plan.strips, 2);
    const int n_int = plan.strips - plan.n_edge_strips;
    const bool both_edges = (g.grow0 == 0 && g.grow0 + g.rows == g.grows);
    const long slots = (long)occ_cache[wpc] * sm_count * wpc / b.n;  // warps available per field
    const int min_mult = tune.min_chunk_mult > 0 ? tune.min_chunk_mult : 2;  // profiles/tune_small_r01.log
    const double edge_cost = tune.edge_cost_pct > 0 ? tune.edge_cost_pct / 100.0 : kEdgeStripCost;
    ClassPlan ci = {0, 0, 0}, ce = {0, 0, 0};
    if (tune.chunk_rows > 0) {  // manual override: same height for both classes, no trimming
        ci.chunk_rows = ce.chunk_rows = std::min(tune.chunk_rows, g.rows);
        ci.chunks = ce.chunks = (g.rows + ci.chunk_rows - 1) / ci.chunk_rows;
    } else {
        // smallest common cost whose warp count fits the resident slots
        double lo = 4.0 * T, hi = (double)(g.rows + 2 * T) * edge_cost + 1.0;
        for (int it = 0; it < 40; ++it) {
            const double mid = 0.5 * (lo + hi);
            const ClassPlan a = plan_class(g.rows, T, RS, mid, 1.0, both_edges, min_mult);
            const ClassPlan e = plan_class(g.rows, T, RS, mid, edge_cost, both_edges, min_mult);
            const long warps = (long)n_int * a.chunks + (long)plan.n_edge_strips * e.chunks;
            if (warps <= slots)
                hi = mid;
            else
                lo = mid;
        }
        ci = plan_class(g.rows, T, RS, hi, 1.0, both_edges, min_mult);
        ce = plan_class(g.rows, T, RS, hi, edge_cost, both_edges, min_mult);
    }
    plan.chunks[0] = ci.chunks;
    plan.chunk_rows[0] = ci.chunk_rows;
    plan.edge_trim[0] = ci.trim;
    plan.chunks[1] = ce.chunks;
    plan.chunk_rows[1] = ce.chunk_rows;
    plan.edge_trim[1] = ce.trim;
    plan.warps_int = n_int * ci.chunks;
    const long total_warps = (long)plan.warps_int + (long)plan.n_edge_strips * ce.chunks;
    dim3 grid((unsigned)((total_warps + wpc - 1) / wpc), b.n);
    kern<<<grid, wpc * 32, smem, st>>>(g, b, plan);
    return cudaGetLastError();
}

template <int T, bool RHS_REGS, int MINB>
cudaError_t launch_T(const Geom& g, const RelaxBatch& b, bool diffuse, int divmode, const StreamTuning& tune, int sm_count,
                     cudaStream_t st) {
    if (!diffuse) {
        if (b.f[0].aux != nullptr)  // first pressure pass with the divergence fused in: prev = u, rhs = v
            return launch_one<T, false, F2D_DIV_F64, 2, RHS_REGS, MINB>(g, b, tune, sm_count, st);
        if (b.f[0].prev == nullptr) return launch_one<T, false, F2D_DIV_F64, 1, RHS_REGS, MINB>(g, b, tune, sm_count, st);
        return launch_one<T, false, F2D_DIV_F64, 0, RHS_REGS, MINB>(g, b, tune, sm_count, st);
    }
    if (divmode == F2D_DIV_F64) {
        if (b.f[0].aux != nullptr)  // first diffuse pass with add_sources fused in: prev = field, rhs = source
            return launch_one<T, true, F2D_DIV_F64, 3, RHS_REGS, MINB>(g, b, tune, sm_count, st);
        return launch_one<T, true, F2D_DIV_F64, 0, RHS_REGS, MINB>(g, b, tune, sm_count, st);
    }
    if (b.f[0].aux != nullptr) return launch_one<T, true, F2D_DIV_F32_CORR, 3, RHS_REGS, MINB>(g, b, tune, sm_count, st);
    return launch_one<T, true, F2D_DIV_F32_CORR, 0, RHS_REGS, MINB>(g, b, tune, sm_count, st);
}

}  // namespace

bool stream_supported(const Geom& g, int T) {
    if (T != 1 && T != 2 && T != 4 && T != 8) return false;
    return g.cols >= 4 && (g.cols % 4 == 0) && (g.pitch % 4 == 0) && g.rows >= 3;
}

cudaError_t launch_jacobi_stream(const Geom& g, const RelaxBatch& b, bool diffuse, int divmode, int T, int sweeps,
                                 const StreamTuning& tune, int sm_count, cudaStream_t st) {
    (void)sweeps;  // == T: the step driver decomposes K into passes of 8/4/2/1 sweeps
    // T = 8 keeps the right-hand side in the smem ring (unroll 3): with a register ring the unrolled
    // row loop (9 x 8 levels) outgrows the instruction cache.  MINB = resident 128-thread CTAs per SM the register
    // allocator must allow: T = 8 -> 3 (168 regs), T = 4 -> 4 (128 regs).
    switch (T) {
        case 1: return launch_T<1, true, 6>(g, b, diffuse, divmode, tune, sm_count, st);
        case 2: return launch_T<2, true, 6>(g, b, diffuse, divmode, tune, sm_count, st);
        case 4: return launch_T<4, true, 4>(g, b, diffuse, divmode, tune, sm_count, st);
        default: return launch_T<8, false, 3>(g, b, diffuse, divmode, tune, sm_count, st);
    }
}

}  // namespace f2d
