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
//   * The row loop is unrolled by RS (a multiple of 3) so that all window/ring register indices
//     are compile-time constants; a row step is one basic block; blocks of RS rows in which no
//     level meets a GLOBAL edge row take a FAST path without range or edge-row checks.
//   * Two fused first passes (template parameter PIN_ZERO): == 2, the first pressure pass computes
//     the divergence from u and v on the fly (p0 == 0) and writes it out for the later passes;
//     == 3, the first diffuse pass forms x0 = FMA(dt, source, field) (add_sources), relaxes from it
//     and writes it out as the right-hand side of the later passes.
//   * A host planner (launch_one) sizes the chunks from a cost model so that one launch is exactly
//     one wave of resident warps that finish together.
//   * The kernel is bound by instruction issue (profiles/ncu_jacobi_*_T8_r02_v5.md), so everything that is not one
//     of the 4 (pressure: levels carried as 4^s * p, see relax_row) / 6 (diffuse: division by the constant 1 + 4a as
//     a correctly rounded two-operation multiplication, f2d_common.cuh) floating-point operations per cell is kept out
//     of the row loop: the edge-column fix exists only in the variant run by the two strips that hold a domain edge
//     column (template parameter EDGE, warp-uniform choice outside the loop); global rows are addressed with running
//     32-bit element offsets (one add per step); no register copy of the output row is kept (the one edge rule that
//     needs the previous output row reads it back); with the right-hand side in shared memory (T = 8) its ring is a
//     plain 16-slot ring, 8 KB-aligned (F2D_RHS_MIRROR=1: the mirrored ring of the first round-2 version, rows written
//     twice so that every level reads at a compile-time offset -- slower since the instruction diet).
//   * At T = 8 the levels 4..7 run one row behind the levels 0..3 (level_lag): two independent dependency chains per
//     row step instead of one.
//   * The passes are launched with programmatic dependent launch (StreamTuning::pdl): a pass is scheduled and set up
//     while its predecessor drains and waits with griddepcontrol.wait before it touches global memory.
//   * A packed fp32x2 variant (FADD2 / FMUL2 / FFMA2, register pairs (x,z) / (y,w)) was built in round 2 and is
//     bit-identical too, but slower: 72.4 vs 59.1 us per pressure pass, 89 vs 81 us per diffuse pass under ncu at
//     4096^2 -- the packed operations occupy the FMA pipe for two cycles each (no pipe time saved,
//     profiles/ubench_fp32x2_r02.jsonl), pair assembly costs two moves per row and level, and the 168-register
//     budget spills (git 7a65d0d; profiles/ncu_jacobi_*_T8_r02_packed_fp32x2.md).
//   * Also built, measured and dropped in round 2 (git history; numbers in profiles/ab_r02_*.log): (a) the passes of one
//     relaxation chained inside ONE launch, each warp waiting only for the progress counters of its 3 x 3 neighbouring
//     (strip, chunk) warps: the device-scope fence + counter round trip per pass costs as much as the launch it replaces
//     inside a CUDA graph (4096^2 step 3.44 vs 3.37 ms, 256^2 0.148 vs 0.109 ms); (b) "rhs generations": each rhs row read
//     back from shared memory once per three levels and kept in registers for the two following steps (3 or 4 instead of
//     8 LDS.128 per row step): shared-memory wavefronts fall from 10.3 M to 6.6 M per pass, but the 24-36 registers that
//     stay live across steps spill, and with 3 x 70 KB of shared memory per SM the L1 left for local memory is tiny
//     (long_scoreboard 1.4 cycles per instruction): pressure pass 63-70 us against 53 us.
#include <algorithm>

#include "f2d_kernels.cuh"

// A/B switches (tools/build_variants.sh builds one library per combination)
#ifndef F2D_LEVEL_SPLIT
#define F2D_LEVEL_SPLIT 1  // T = 8: levels 4..7 run one row behind levels 0..3 (see level_lag)
#endif
#ifndef F2D_PRESSURE_SCALED
#define F2D_PRESSURE_SCALED 1  // pressure levels carried as 4^s * p (see relax_row): 4 instead of 5 operations per cell-sweep
#endif
#ifndef F2D_RHS_MIRROR
#define F2D_RHS_MIRROR 0  // rhs ring in shared memory: 1 = 16 slots + mirror of slots 0..7 (every row of a step at a
                          // compile-time offset below one pointer, rows with bit 3 clear are written twice); 0 = plain
                          // 16-slot ring, 8 KB-aligned, one "(row << 9) & mask | base" per read
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
// rhs smem ring: register mode only needs the landing zone (kRingP slots); smem mode keeps rows r+PFD .. r-T in
// kRingR slots and mirrors slots 0 .. kMirror-1 behind them (see slot_rd)
constexpr int kRingR = 16;   // >= PFD + T + 2 for T <= 8
constexpr int kMirror = 8;   // >= T
__host__ __device__ constexpr int ring_r_slots(bool rhs_regs) { return rhs_regs ? kRingP : kRingR + (F2D_RHS_MIRROR ? kMirror : 0); }

// Level chain.  Without a lag, level s+1 consumes in the same row step the row level s has just produced: the T levels
// of a step are ONE dependent chain (FADD -> ... -> next level), and a warp can only overlap the four cells of its
// float4.  With F2D_LEVEL_SPLIT the upper half of the levels (T = 8: levels 4..7) works one row behind: level 4 reads
// the row level 3 produced in the PREVIOUS step, so a step consists of two independent chains of four levels that the
// scheduler interleaves.  Cost: one more drain row per chunk and one more rhs row in the ring; no extra registers
// across steps (level 3's new row replaces the row level 4 consumed last, in the same window slot).
__host__ __device__ constexpr int level_split(int T, bool rhs_regs) { return (F2D_LEVEL_SPLIT && T == 8 && !rhs_regs) ? 4 : T; }
__host__ __device__ constexpr int level_lag(int T, bool rhs_regs, int s) { return s >= level_split(T, rhs_regs) ? 1 : 0; }
__host__ __device__ constexpr int max_lag(int T, bool rhs_regs) { return level_lag(T, rhs_regs, T - 1); }
static_assert(kPFD + 8 + max_lag(8, false) + 1 <= kRingR, "rhs ring: rows r + PFD .. r - T - lag");

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
__device__ __forceinline__ float4 ld_global_f4(const float* p) {
    float4 v;
    asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_global_f4(float* p, const float4& v) {
    asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// Pressure sweep (gpu.cu:187-188): p' = ((((d + pE) + pW) + pS) + pN) * 0.25f.  The multiplication by 0.25 is exact, so
// a pass can carry level s as t_s = 4^s * p_s:  t_{s+1} = (((FMA(d, 4^s, tE) + tW) + tS) + tN)  -- FMA(d, 4^s, tE) rounds
// 4^s * (d + pE) exactly as the reference rounds d + pE, and every later sum is the reference's sum times 4^s -- and
// multiply the last level by 4^-T once.  4T + 1 instead of 5T operations per cell and pass, bit-identical to the
// reference unless 0.25 * sum is subnormal (|p| < 2^-126: the reference then rounds to the subnormal grid at every
// sweep, this only at the end of the pass) or 4^T * p overflows (|p| > 2^111).
__host__ __device__ constexpr float level_scale(int s) { return (float)(1u << (2 * s)); }
__host__ __device__ constexpr float pass_unscale(int T) { return 1.0f / (float)(1u << (2 * T)); }
__device__ __forceinline__ float pressure_scaled(float dv, float cs, float e, float w, float s, float n) {
    return __fadd_rn(__fadd_rn(__fadd_rn(__fmaf_rn(dv, cs, e), w), s), n);
}

// level S of a pass of T: rows a (north), b (centre), c (south) of level S -> the centre row of level S + 1
// (S is a constant after unrolling, so the scale factors fold into immediates)
template <bool DIFFUSE, int DIVMODE, int T>
__device__ __forceinline__ float4 relax_row(const float4& a, const float4& b, const float4& c, float l, float rt,
                                            const float4& rhs, const DiffuseCoef& k, int S) {
    float4 o;
    if (DIFFUSE) {
        o.x = diffuse_update<DIVMODE>(l, b.y, a.x, c.x, rhs.x, k);
        o.y = diffuse_update<DIVMODE>(b.x, b.z, a.y, c.y, rhs.y, k);
        o.z = diffuse_update<DIVMODE>(b.y, b.w, a.z, c.z, rhs.z, k);
        o.w = diffuse_update<DIVMODE>(b.z, rt, a.w, c.w, rhs.w, k);
    } else if (F2D_PRESSURE_SCALED) {
        const float cs = level_scale(S);
        o.x = pressure_scaled(rhs.x, cs, b.y, l, c.x, a.x);
        o.y = pressure_scaled(rhs.y, cs, b.z, b.x, c.y, a.y);
        o.z = pressure_scaled(rhs.z, cs, b.w, b.y, c.z, a.z);
        o.w = pressure_scaled(rhs.w, cs, rt, b.z, c.w, a.w);
        if (S == T - 1) {
            const float un = pass_unscale(T);
            o.x = __fmul_rn(o.x, un);
            o.y = __fmul_rn(o.y, un);
            o.z = __fmul_rn(o.z, un);
            o.w = __fmul_rn(o.w, un);
        }
    } else {
        o.x = pressure_update(rhs.x, b.y, l, c.x, a.x);
        o.y = pressure_update(rhs.y, b.z, b.x, c.y, a.y);
        o.z = pressure_update(rhs.z, b.w, b.y, c.z, a.z);
        o.w = pressure_update(rhs.w, rt, b.z, c.w, a.w);
    }
    return o;
}
// factor that takes a corner value carried at the scale of level S to the scale of the row level S produces
template <bool DIFFUSE, int T>
__host__ __device__ constexpr float corner_carry(int S) {
    return (DIFFUSE || !F2D_PRESSURE_SCALED) ? 1.0f : (S == T - 1 ? 1.0f / level_scale(S) : 4.0f);
}

// edge row from the adjacent interior row of the same level; corner cells keep `keep`
__device__ __forceinline__ float4 edge_row(const float4& inner, const float4& keep, float keep_mul, bool neg, bool has_left,
                                           bool has_right) {
    float4 o;
    o.x = apply_sign(inner.x, neg);
    o.y = apply_sign(inner.y, neg);
    o.z = apply_sign(inner.z, neg);
    o.w = apply_sign(inner.w, neg);
    if (has_left) o.x = __fmul_rn(keep.x, keep_mul);  // exact: keep_mul is a power of four (1 for diffuse)
    if (has_right) o.w = __fmul_rn(keep.w, keep_mul);
    return o;
}

// ---- shared-memory rings (32-bit shared addresses, lane offset included in the base)
// landing rings (kRingP slots of one row each, the block aligned to its own size): slot of row `row`
__device__ __forceinline__ unsigned slot8(unsigned base, int row) { return (((unsigned)row << 9) & ((kRingP - 1) << 9)) | base; }
// Mirrored rhs ring (rhs in shared memory): row i lives in slot i & 15; rows with (i & 8) == 0 are ALSO written to
// slot 16 + (i & 15).  The rows r-1 .. r-T (T <= 8) a step reads are then contiguous below one read pointer:
//   (r & 8) != 0: slots (r & 15) - 1 ... >= 0, the plain ring;
//   (r & 8) == 0: from 16 + (r & 15) downwards -- mirrors first, then the plain slots 15, 14, ...
// so level s reads its row at the compile-time offset -(s + 1) * 512 from slot_rd(base, r).
#if F2D_RHS_MIRROR
__device__ __forceinline__ unsigned slot_wr(unsigned base, int row) { return base + (((unsigned)row << 9) & ((kRingR - 1) << 9)); }
__device__ __forceinline__ bool mirrored(int row) { return (row & kMirror) == 0; }
__device__ __forceinline__ unsigned slot_rd(unsigned base, int r) { return slot_wr(base, r) + (mirrored(r) ? (unsigned)(kRingR << 9) : 0u); }
// rhs row r - back (1 <= back <= T) given rd = slot_rd(base, r)
__device__ __forceinline__ unsigned slot_back(unsigned, unsigned rd, int, int back) { return rd - (unsigned)(back << 9); }
#else
// plain ring, the block aligned to its own size (8 KB)
__device__ __forceinline__ unsigned slot_wr(unsigned base, int row) { return (((unsigned)row << 9) & ((kRingR - 1) << 9)) | base; }
__device__ __forceinline__ bool mirrored(int) { return false; }
__device__ __forceinline__ unsigned slot_rd(unsigned, int r) { return (unsigned)r << 9; }
__device__ __forceinline__ unsigned slot_back(unsigned base, unsigned rd, int, int back) {
    return ((rd - (unsigned)(back << 9)) & ((kRingR - 1) << 9)) | base;
}
#endif
static_assert(kMirror == 8 && kRingR == 16, "the mirror rule is bit 3 of the row index");

// per-warp constants of one (strip, chunk)
struct Ctx {
    const float* prev;  // lane-adjusted: + column of this lane
    const float* rhs;
    float* next;
    unsigned sp, sr;  // the two async-copy rings (sp: iterate, always a landing ring; sr: rhs -- a landing ring, or the
                      // mirrored ring when the relaxation reads its rhs from shared memory)
    unsigned sd;      // fused first passes, rhs in shared memory: mirrored ring of the COMPUTED rhs rows
    float* aux;       // fused divergence: the divergence field written for the later passes; fused add_sources: x0
    float mhalf_h;    // fused divergence: -0.5f * h; fused add_sources: dt
    int row_top, row_bot;  // local index of the global top / bottom edge row (or out of range)
    DiffuseCoef coef;
    int pitch, rs, re, y0, y1;
    int cp_bytes;
    bool top_dom, bot_dom, own_x, has_left, has_right, edge_warp, neg_c, neg_r;
};

// running element offsets of the rows a step touches (advanced by one pitch per step: no per-row multiplication)
struct Run {
    int off_in;   // row r + PFD (the row whose async copy is issued)
    int off_out;  // row r - T   (the row the last level stores)
};

// RS consecutive row steps starting at relative row rb (a multiple of RS, so rb % 3 == 0 and the
// register slots of every row are compile-time constants).
//
// FAST blocks run every level unconditionally.  During the warm-up / drain of a chunk some levels
// then work on rows outside the chunk's input range (garbage, but finite): those results only ever
// feed cells outside the dependency cone of the rows this warp stores, and the store itself is
// predicated on the owned row range.  Only blocks in which a level meets a GLOBAL top or bottom edge
// row (edge rule, corner carry) take the checked path (FAST == false).
// EDGE: this warp's strip holds a domain edge column; interior strips are compiled without the edge-column fix.
template <int T, bool DIFFUSE, int DIVMODE, int PIN_ZERO, bool RHS_REGS, bool FAST, bool EDGE, int RS, int NRH>
__device__ __forceinline__ void run_block(const Ctx& cx, Run& st, int rb, int nsteps, float4 (&W)[T][3], float4 (&RH)[NRH],
                                          float (&wl)[T], float (&er)[T], float4 (&UV)[2][3]) {
    static_assert(T <= kMirror, "the mirrored ring covers T <= 8 rows");
    // PIN_ZERO == 2: the first pressure pass with the divergence fused in (gpu.cu:164-177 + :376): the two
    // async rings carry u and v rows instead of iterate and rhs; the rhs row r-1 is computed on the fly
    constexpr bool FUSE = (PIN_ZERO == 2);
    // PIN_ZERO == 3: the first diffuse pass with add_sources fused in (gpu.cu:56-67 folded into :290-312): the
    // rings carry the field and its source; x0 = FMA(dt, s, f) is formed per row, used as iterate AND rhs,
    // and written out as the rhs of the later passes
    constexpr bool FSRC = (PIN_ZERO == 3);
    constexpr bool RHS_LANDS_MIRRORED = !RHS_REGS && !FUSE && !FSRC;  // the async copy itself fills the mirrored ring
    constexpr int SPLIT = level_split(T, RHS_REGS), LAG = max_lag(T, RHS_REGS);  // st.off_out follows row r - T - LAG
    const bool hl = EDGE && cx.has_left, hr = EDGE && cx.has_right;
#pragma unroll
    for (int k = 0; k < RS; ++k) {
        const int rr = rb + k;
        if (rr >= nsteps) break;
        const int r = cx.rs + rr;

        // 1. keep PFD rows in flight
        {
            const int rl = r + kPFD;
            if (rl <= cx.re) {
                if (PIN_ZERO != 1) cp_async16_s(slot8(cx.sp, rl), cx.prev + st.off_in, cx.cp_bytes);
                if (RHS_LANDS_MIRRORED) {
                    // two copies, branch-free: the slot and its mirror (a row without a mirror is simply written to its
                    // slot twice; a branch here would split the row step into several basic blocks)
                    const unsigned w = slot_wr(cx.sr, rl);
                    cp_async16_s(w, cx.rhs + st.off_in, cx.cp_bytes);
#if F2D_RHS_MIRROR
                    cp_async16_s(w + (mirrored(rl) ? (unsigned)(kRingR << 9) : 0u), cx.rhs + st.off_in, cx.cp_bytes);
#endif
                } else {
                    cp_async16_s(slot8(cx.sr, rl), cx.rhs + st.off_in, cx.cp_bytes);
                }
            }
            cp_async_commit();
        }
        // 2. row r has landed (each lane reads back only the 16 bytes it copied itself)
        cp_async_wait<kPFD>();
        if (FAST || r <= cx.re) {  // FAST: past the last input row this re-reads a stale ring slot (harmless)
            if (FSRC) {
                const float4 f4 = lds128(slot8(cx.sp, r));
                const float4 s4 = lds128(slot8(cx.sr, r));
                const bool row_in = (r != cx.row_top) && (r != cx.row_bot);  // global interior row
                float4 x0;
                x0.x = (row_in && !hl) ? __fmaf_rn(cx.mhalf_h, s4.x, f4.x) : f4.x;  // mhalf_h carries dt here
                x0.y = row_in ? __fmaf_rn(cx.mhalf_h, s4.y, f4.y) : f4.y;
                x0.z = row_in ? __fmaf_rn(cx.mhalf_h, s4.z, f4.z) : f4.z;
                x0.w = (row_in && !hr) ? __fmaf_rn(cx.mhalf_h, s4.w, f4.w) : f4.w;
                W[0][m3(k)] = x0;
                if (RHS_REGS) {
                    RH[mrs(k, RS)] = x0;
                } else {
                    const unsigned w = slot_wr(cx.sd, r);
                    sts128(w, x0);
                    if (mirrored(r)) sts128(w + (kRingR << 9), x0);
                }
                if (cx.own_x && r >= cx.y0 && r < cx.y1 && r <= cx.re) st_global_f4(cx.aux + (st.off_out + (T + LAG) * cx.pitch), x0);
            } else if (!PIN_ZERO)
                W[0][m3(k)] = lds128(slot8(cx.sp, r));
            else
                W[0][m3(k)] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (FSRC) {
            } else if (FUSE) {
                UV[0][m3(k)] = lds128(slot8(cx.sp, r));  // u row r
                UV[1][m3(k)] = lds128(slot8(cx.sr, r));  // v row r
            } else if (RHS_REGS) {
                RH[mrs(k, RS)] = lds128(slot8(cx.sr, r));
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
            dv.x = hl ? dv.y : dv.x;
            dv.w = hr ? dv.z : dv.w;
            const int qd = r - 1;
            if (RHS_REGS) {
                RH[mrs(k - 1, RS)] = dv;
            } else {
                const unsigned w = slot_wr(cx.sd, qd);
                sts128(w, dv);
                if (mirrored(qd)) sts128(w + (kRingR << 9), dv);
            }
            if (cx.own_x && qd >= cx.rs + 1 && qd <= cx.re - 1) {
                if (qd >= cx.y0 && qd < cx.y1) st_global_f4(cx.aux + (st.off_out + (T + LAG - 1) * cx.pitch), dv);
                // the edge rows of the stored field copy the adjacent interior row, corners are the memset zeros
                float4 e = dv;
                e.x = hl ? 0.f : e.x;
                e.w = hr ? 0.f : e.w;
                if (qd == 1 && cx.top_dom && cx.y0 == 0) st_global_f4(cx.aux, e);
                if (qd == cx.re - 1 && cx.bot_dom) st_global_f4(cx.aux + (size_t)cx.re * cx.pitch, e);
            }
        }
        // west/east neighbours of the centre rows of all levels (rows produced in earlier steps)
#pragma unroll
        for (int s = 0; s < T; ++s) {
            const float4 bb = W[s][m3(k - s - 1 - level_lag(T, RHS_REGS, s))];
            wl[s] = __shfl_up_sync(0xffffffffu, bb.w, 1);
            er[s] = __shfl_down_sync(0xffffffffu, bb.x, 1);
        }

        // 3. level s+1 produces row q = r - s - 1 - lag(s) from level s rows q-1, q, q+1; wl[s], er[s] are the
        //    west/east neighbours of the centre row.  A level is active while rs+1 <= q <= re-1.
        //    The lagging levels go FIRST in program order: level SPLIT still reads the window slot that level SPLIT-1
        //    overwrites in this step (rows q-1 of the one and q of the other share a slot).
        const unsigned rbase = (FUSE || FSRC) ? cx.sd : cx.sr;
        const unsigned rd = RHS_REGS ? 0u : slot_rd(rbase, r);
#pragma unroll
        for (int si = 0; si < T; ++si) {
            const int s = (si < T - SPLIT) ? SPLIT + si : si - (T - SPLIT);
            const int lag = level_lag(T, RHS_REGS, s);
            const int q = r - s - 1 - lag;
            const int sa = m3(k - s - 2 - lag), sm = m3(k - s - 1 - lag), sc = m3(k - s - lag);
            const int sn = (s + 1 < T) ? s + 1 : 0;  // keeps the dead branch's index in range
            if (FAST || (q >= cx.rs + 1 && q <= cx.re - 1)) {
                const float4 a = W[s][sa], b = W[s][sm], c = W[s][sc];
                const float l = wl[s], rt = er[s];
                float4 rhs;
                if (RHS_REGS)
                    rhs = RH[mrs(k - s - 1, RS)];
                else
                    rhs = lds128(slot_back(rbase, rd, r, s + 1 + lag));
                float4 nw = relax_row<DIFFUSE, DIVMODE, T>(a, b, c, l, rt, rhs, cx.coef, s);
                // interior rows of an edge strip: col 0 = +/- col 1, col N-1 = +/- col N-2 (gpu.cu:16-17, 31-32), as
                // two predicated selects (no branch) so that a whole row step stays one basic block
                if (EDGE) {
                    nw.x = cx.has_left ? apply_sign(nw.y, cx.neg_c) : nw.x;
                    nw.w = cx.has_right ? apply_sign(nw.z, cx.neg_c) : nw.w;
                }
                if (s + 1 < T) {
                    W[sn][sm] = nw;
                } else {
                    if (cx.own_x && q >= cx.y0 && q < cx.y1) st_global_f4(cx.next + st.off_out, nw);
                }
                if (!FAST && cx.top_dom && q == 1) {  // global top edge row of the same level (corners kept)
                    const float4 e = edge_row(nw, a, corner_carry<DIFFUSE, T>(s), cx.neg_r, hl, hr);
                    if (s + 1 < T)
                        W[sn][sa] = e;
                    else if (cx.own_x && cx.y0 == 0)
                        st_global_f4(cx.next, e);
                }
            } else if (!FAST && cx.bot_dom && q == cx.re && q >= cx.rs + 1) {  // global bottom edge row
                // the adjacent interior row of the last level was stored one step ago by this very lane (the bottom chunk
                // owns it): read it back instead of keeping every output row alive in registers for this one step
                const float4 inner = (s + 1 < T) ? W[sn][sa] : ld_global_f4(cx.next + (st.off_out - cx.pitch));
                const float4 e = edge_row(inner, W[s][sm], corner_carry<DIFFUSE, T>(s), cx.neg_r, hl, hr);
                if (s + 1 < T)
                    W[sn][sm] = e;
                else if (cx.own_x && q >= cx.y0 && q < cx.y1)
                    st_global_f4(cx.next + st.off_out, e);
            }
        }
        st.off_in += cx.pitch;
        st.off_out += cx.pitch;
    }
}

// all row steps of one warp; EDGE is the warp's class (chosen once, outside the row loop)
template <int T, bool DIFFUSE, int DIVMODE, int PIN_ZERO, bool RHS_REGS, bool EDGE, int RS, int NRH>
__device__ __forceinline__ void march(const Ctx& cx, Run& st, int nsteps, int top_lo, int top_hi, int bot_lo, int bot_hi,
                                      float4 (&W)[T][3], float4 (&RH)[NRH], float (&wl)[T], float (&er)[T],
                                      float4 (&UV)[2][3]) {
    for (int rb = 0; rb < nsteps; rb += RS) {
        const int r_first = cx.rs + rb, r_last = r_first + RS - 1;
        const bool edge_block = (r_first <= top_hi && r_last >= top_lo) || (r_first <= bot_hi && r_last >= bot_lo);
        if (!edge_block)
            run_block<T, DIFFUSE, DIVMODE, PIN_ZERO, RHS_REGS, true, EDGE, RS, NRH>(cx, st, rb, nsteps, W, RH, wl, er, UV);
        else
            run_block<T, DIFFUSE, DIVMODE, PIN_ZERO, RHS_REGS, false, EDGE, RS, NRH>(cx, st, rb, nsteps, W, RH, wl, er, UV);
    }
}

// MINB = resident CTAs (of 128 threads) per SM the register allocator must allow
template <int T, bool DIFFUSE, int DIVMODE, int PIN_ZERO, bool RHS_REGS, int MINB>
__global__ void __launch_bounds__(128, MINB) k_jacobi_stream(Geom g, RelaxBatch batch, StreamPlan plan) {
    constexpr int HALO = halo_of(T);
    constexpr int RS = RHS_REGS ? rs_of(T) : 3;  // unroll factor of the row loop
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
    cx.coef = fld.coef;
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
        // [rhs rings of warp 0 .. wpc-1][landing rings ...]: a landing ring is 4 KB and 4 KB-aligned ("(row << 9) & mask
        // | base" addresses a slot with two integer instructions); the mirrored rhs ring is 12 KB per warp
        constexpr unsigned RB = ring_r_slots(RHS_REGS) * kLanes * 16u, PB = kRingP * kLanes * 16u;
        constexpr unsigned AL = (!F2D_RHS_MIRROR && !RHS_REGS) ? RB : PB;  // a plain rhs ring is aligned to its own size
        static_assert(AL >= PB && (AL & (AL - 1u)) == 0u, "ring alignment");
        const unsigned s0 = ((unsigned)__cvta_generic_to_shared(smem) + AL - 1u) & ~(AL - 1u);
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
    // the chunk that owns the bottom edge row also owns the interior row above it: the edge rule of the last level reads
    // that row back from what this warp stored one step earlier (run_block)
    if (chunk == n_chunks - 1 && chunk > 0) cx.y0 = min(cx.y0, g.rows - 2);
    if (chunk == n_chunks - 2) cx.y1 = min(cx.y1, g.rows - 2);
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
    constexpr int LAG = max_lag(T, RHS_REGS);
    const int nsteps = (cx.y1 - 1 + T + LAG) - cx.rs + 1;
    // absolute input rows r during which some level meets a global edge row:
    //   top:    q == 1 at level s+1  <=>  r = s + 2 + lag(s),      s in [0, T)
    //   bottom: q == re at level s+1 <=>  r = re + s + 1 + lag(s), s in [0, T)
    const int top_lo = cx.top_dom ? 2 : 1 << 30, top_hi = cx.top_dom ? T + 1 + LAG : -1;
    const int bot_lo = cx.bot_dom ? cx.re + 1 : 1 << 30, bot_hi = cx.bot_dom ? cx.re + T + LAG : -1;

    float4 W[T][3];
    float4 RH[NRH];
    float wl[T], er[T];
    float4 UV[2][3];
#pragma unroll
    for (int m = 0; m < 3; ++m) UV[0][m] = UV[1][m] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int s = 0; s < T; ++s) {
        wl[s] = er[s] = 0.f;
#pragma unroll
        for (int m = 0; m < 3; ++m) W[s][m] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int m = 0; m < NRH; ++m) RH[m] = make_float4(0.f, 0.f, 0.f, 0.f);

    // programmatic dependent launch (StreamTuning::pdl): everything above ran while the previous kernel of the stream was
    // still draining; its results (and the buffers it read, which this pass overwrites) are safe to touch from here on.
    // The next pass may be scheduled as soon as CTAs of this one retire.  Both are no-ops in a normal launch.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    // ---- prologue: rows rs .. rs+PFD-1 in flight
    constexpr bool RHS_LANDS_MIRRORED = !RHS_REGS && PIN_ZERO < 2;
#pragma unroll
    for (int p = 0; p < kPFD; ++p) {
        const int rl = cx.rs + p;
        if (rl <= cx.re) {
            const size_t off = (size_t)rl * cx.pitch;
            if (PIN_ZERO != 1) cp_async16_s(slot8(cx.sp, rl), cx.prev + off, cx.cp_bytes);
            if (RHS_LANDS_MIRRORED) {
                const unsigned w = slot_wr(cx.sr, rl);
                cp_async16_s(w, cx.rhs + off, cx.cp_bytes);
                if (mirrored(rl)) cp_async16_s(w + (kRingR << 9), cx.rhs + off, cx.cp_bytes);
            } else {
                cp_async16_s(slot8(cx.sr, rl), cx.rhs + off, cx.cp_bytes);
            }
        }
        cp_async_commit();
    }
    Run st;
    st.off_in = (cx.rs + kPFD) * cx.pitch;
    st.off_out = (cx.rs - T - LAG) * cx.pitch;

    if (!cx.edge_warp)
        march<T, DIFFUSE, DIVMODE, PIN_ZERO, RHS_REGS, false, RS, NRH>(cx, st, nsteps, top_lo, top_hi, bot_lo, bot_hi, W, RH, wl, er, UV);
    else
        march<T, DIFFUSE, DIVMODE, PIN_ZERO, RHS_REGS, true, RS, NRH>(cx, st, nsteps, top_lo, top_hi, bot_lo, bot_hi, W, RH, wl, er, UV);
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
    // rhs ring(s) + iterate landing ring(s) per warp (see the kernel's layout comment) + 4 KB of alignment slack
    const size_t smem = (size_t)wpc * (ring_r_slots(RHS_REGS) + kRingP + (PIN_ZERO >= 2 ? kRingP : 0)) * kLanes * sizeof(float4) +
                        (size_t)((!F2D_RHS_MIRROR && !RHS_REGS) ? kRingR : kRingP) * kLanes * sizeof(float4);
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
    if (tune.pdl) {
        cudaLaunchConfig_t lc = {};
        lc.gridDim = grid;
        lc.blockDim = dim3(wpc * 32);
        lc.dynamicSmemBytes = smem;
        lc.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        lc.attrs = at;
        lc.numAttrs = 1;
        return cudaLaunchKernelEx(&lc, kern, g, b, plan);
    }
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
