// f2d_gs_tile.h -- tile core of the Gauss-Seidel wavefront (the fluid_solver_cpu-compatible relaxation).
//
// fluid_solver_cpu relaxes IN PLACE in lexicographic order (src/fluid_solver_cpu.cpp:104-113 diffuse,
// :196-204 pressure): cell (i,j) of sweep k reads the sweep-k values of (i-1,j), (i,j-1) and the
// sweep-(k-1) values of (i+1,j), (i,j+1), then set_boundary runs (cpp:112, :203).  Any execution order that
// respects those four dependencies produces the same bits, so the grid is cut into tiles of kBand rows x
// kTileCols columns and tile T(k, w, c) (sweep k, row band w, column tile c) may run as soon as
//     T(k, w-1, c)   wrote the row above it        (new north values),
//     T(k, w, c-1)   wrote the column left of it    (new west values; the same warp's previous tile),
//     T(k-1, w+1, c) wrote the row below it         (old south values),
//     T(k-1, w, c+1) wrote the column right of it   (old east values; implies T(k-1, w, c)),
// i.e. the tiles of ALL sweeps form one wavefront  tau = w + c + 2k  and run concurrently on one in-place
// array.  The same four conditions also cover the write-after-read hazards (a tile overwrites sweep-(k-1)
// values only after every reader of them has finished).  Inside a tile the 32 lanes of a warp own one row
// each and march along the columns skewed by one step per row (lane l handles column t - l at step t).
//
// Edge cells belong to the tile whose interior cells they mirror: a tile that touches row 1 / row R-2 /
// column 1 / column C-2 also writes row 0 / R-1 / column 0 / C-1 with the sign of the boundary kind
// (cpp:33-83 without the corners), so the next sweep reads them like any other neighbour.  The four corners
// are never read by a 5-point stencil; they are averaged once after the last sweep (k_corners_avg).
//
// This header is plain C++ on purpose: the CUDA kernel (f2d_gs.cu) calls the functions below with
// lane = threadIdx.x % 32 and __syncwarp() between steps, and tests/gs_emul.cpp compiles the SAME functions
// with g++ and runs the tiles in random dependency-respecting orders against the sequential sweep (no GPU needed).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define F2D_HD __host__ __device__ __forceinline__
#else
#define F2D_HD inline
#endif

namespace f2d {
namespace gs {

constexpr int kBand = 32;                          // rows per band == lanes of a warp
constexpr int kTileCols = 32;                      // interior columns per tile
constexpr int kTP = kTileCols + 2;                 // pitch of the value tile; kTP - 1 is odd: skewed accesses hit 32 banks
constexpr int kTileFloats = (kBand + 2) * kTP;     // value tile with a one-cell frame
constexpr int kRhsFloats = kBand * kTileCols;      // right-hand side tile (pitch kTileCols: 31 * l + t, conflict-free)
// One warp's on-chip block: [kPad slack][value tile][rhs tile][kPad slack], contiguous.  The step loop addresses its
// operands with unclamped column offsets; a lane whose column is outside the tile (the skew's ramp-up / ramp-down)
// then reads up to 31 floats before or after its row.  Those reads are discarded, the slack keeps them in bounds.
#ifndef F2D_GS_PAD
#define F2D_GS_PAD 64
#endif
constexpr int kPad = F2D_GS_PAD;  // >= 32
constexpr int kWarpFloats = kPad + kTileFloats + kRhsFloats + kPad;

// boundary kinds as in include/f2d.h
constexpr int kBndContinuous = 0, kBndOppositeHorizontal = 1, kBndOppositeVertical = 2;

struct Shape {
    int rows, cols, pitch;  // field extent and row pitch in floats
    int nb, nt;             // row bands, column tiles
};

F2D_HD Shape make_shape(int rows, int cols, int pitch) {
    Shape s;
    s.rows = rows;
    s.cols = cols;
    s.pitch = pitch;
    s.nb = (rows - 2 + kBand - 1) / kBand;
    s.nt = (cols - 2 + kTileCols - 1) / kTileCols;
    return s;
}

struct Tile {
    int i0, nr;  // first interior row of the band, rows in it (1..kBand)
    int j0, nc;  // first interior column of the tile, columns in it (1..kTileCols)
};

F2D_HD Tile make_tile(const Shape& s, int w, int c) {
    Tile t;
    t.i0 = 1 + kBand * w;
    t.nr = s.rows - 1 - t.i0;
    if (t.nr > kBand) t.nr = kBand;
    t.j0 = 1 + kTileCols * c;
    t.nc = s.cols - 1 - t.j0;
    if (t.nc > kTileCols) t.nc = kTileCols;
    return t;
}

// ---- progress flags ---------------------------------------------------------------------------------
// done[(k * nb + w)] = number of column tiles band w has finished in sweep k (one block of K * nb words per
// problem).  Before tile c the warp of (k, w) waits for up to three counters:
struct Deps {
    int n;
    int idx[3];        // index into the problem's flag block
    unsigned need[3];  // minimal value
};

F2D_HD Deps tile_deps(const Shape& s, int k, int w, int c) {
    Deps d;
    d.n = 0;
    if (w > 0) {  // north row: T(k, w-1, c)
        d.idx[d.n] = k * s.nb + (w - 1);
        d.need[d.n++] = (unsigned)(c + 1);
    }
    if (k > 0) {
        if (w + 1 < s.nb) {  // south row: T(k-1, w+1, c)
            d.idx[d.n] = (k - 1) * s.nb + (w + 1);
            d.need[d.n++] = (unsigned)(c + 1);
        }
        // east column: T(k-1, w, c+1); the last tile only needs its own previous sweep
        d.idx[d.n] = (k - 1) * s.nb + w;
        d.need[d.n++] = (unsigned)((c + 2 < s.nt) ? c + 2 : s.nt);
    }
    return d;
}

// ---- arithmetic (g++: -ffp-contract=off; nvcc: intrinsics, the library is built with -fmad=false anyway) ----
#if defined(__CUDA_ARCH__)
F2D_HD float fadd(float a, float b) { return __fadd_rn(a, b); }
F2D_HD float fmul(float a, float b) { return __fmul_rn(a, b); }
F2D_HD float fdiv(float a, float b) { return __fdiv_rn(a, b); }
F2D_HD float ld_iter(const float* p) { return __ldcg(p); }  // iterate: written by other SMs, read at L2
F2D_HD float ld_rhs(const float* p) { return __ldg(p); }    // right-hand side: read-only for the whole launch
#else
F2D_HD float fadd(float a, float b) { return a + b; }
F2D_HD float fmul(float a, float b) { return a * b; }
F2D_HD float fdiv(float a, float b) { return a / b; }
F2D_HD float ld_iter(const float* p) { return *p; }
F2D_HD float ld_rhs(const float* p) { return *p; }
#endif

// diffuse (cpp:107-108): (x0 + a * (((N + S) + W) + E)) / (1.f + 4.f * a), product and sum rounded separately
F2D_HD float diffuse_cell(float n, float s, float w, float e, float x0, float a, float c) {
    float sum = fadd(fadd(fadd(n, s), w), e);
    return fdiv(fadd(x0, fmul(a, sum)), c);
}
// pressure (cpp:199-200): ((((div + E) + W) + S) + N) / 4.0f.  Dividing by 4 and multiplying by 0.25 round the same
// real number, so the product is bit-identical (subnormal results included) and skips the divide sequence.
F2D_HD float pressure_cell(float n, float s, float w, float e, float dv) {
    return fmul(fadd(fadd(fadd(fadd(dv, e), w), s), n), 0.25f);
}

// ---- the phases of one tile, per lane -------------------------------------------------------------------
// Phase 1 brings the (nr + 2) x (nc + 2) frame of the iterate and the nr x nc right-hand side on chip.  It is
// split so that the bulk never sits on the critical path:
//   tile_prefetch   the tile's INTERIOR old values and right-hand side -> registers (lane = column, one register
//                   per row).  They are final as soon as T(k-1, w, c) is done, which the wait of tile c-1
//                   already established (its east condition), so the loads are issued before tile c-1 computes
//                   and land while it does;
//   tile_frame_*    after the tile's own wait: the row above, the row below and the column right of the tile
//                   (one batch of loads); the column left of it is the previous tile's last column, still in
//                   shared memory (from global for the first tile);
//   tile_commit     registers -> shared memory.
constexpr int kPf = 2 * kBand;  // prefetch registers per lane

F2D_HD void tile_prefetch(const Shape& s, const Tile& t, const float* x, const float* rhs, float* pf, int lane) {
    // one pointer per array, advanced by the pitch: the address of each of the 64 loads costs one add
    const size_t o = (size_t)t.i0 * s.pitch + t.j0 + lane;
    const float* px = x + o;
    const float* pr = rhs + o;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < kBand; ++r) {
        const bool ok = (r < t.nr) && (lane < t.nc);
        pf[r] = ok ? ld_iter(px) : 0.f;
        pf[kBand + r] = ok ? ld_rhs(pr) : 0.f;
        px += s.pitch;
        pr += s.pitch;
    }
}

struct Frame {
    float top0, top1, bot0, bot1, right, left;
};

// `first` : the band's first tile (left column from global); otherwise `tile` still holds the previous tile.
F2D_HD Frame tile_frame_load(const Shape& s, const Tile& t, const float* x, const float* tile, int lane, bool first) {
    Frame f;
    const float* top = x + (size_t)(t.i0 - 1) * s.pitch + (t.j0 - 1);
    const float* bot = x + (size_t)(t.i0 + t.nr) * s.pitch + (t.j0 - 1);
    const bool hi = (kBand + lane < t.nc + 2);  // lanes 0, 1 also fetch frame columns 32, 33
    f.top0 = (lane < t.nc + 2) ? ld_iter(top + lane) : 0.f;
    f.top1 = hi ? ld_iter(top + kBand + lane) : 0.f;
    f.bot0 = (lane < t.nc + 2) ? ld_iter(bot + lane) : 0.f;
    f.bot1 = hi ? ld_iter(bot + kBand + lane) : 0.f;
    const float* row = x + (size_t)(t.i0 + lane) * s.pitch;
    f.right = (lane < t.nr) ? ld_iter(row + t.j0 + t.nc) : 0.f;
    f.left = 0.f;
    if (lane < t.nr) f.left = first ? ld_iter(row + t.j0 - 1) : tile[(lane + 1) * kTP + kTileCols];
    return f;
}

F2D_HD void tile_commit(const Tile& t, const float* pf, float* tile, float* rt, int lane) {
    if (lane >= t.nc) return;
    float* tc = tile + kTP + lane + 1;
    float* rc = rt + lane;
    const bool full = (t.nr == kBand);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < kBand; ++r)
        if (full || r < t.nr) {
            tc[r * kTP] = pf[r];
            rc[r * kTileCols] = pf[kBand + r];
        }
}

F2D_HD void tile_frame_store(const Tile& t, const Frame& f, float* tile, int lane) {
    const bool hi = (kBand + lane < t.nc + 2);
    if (lane < t.nc + 2) {
        tile[lane] = f.top0;
        tile[(t.nr + 1) * kTP + lane] = f.bot0;
    }
    if (hi) {
        tile[kBand + lane] = f.top1;
        tile[(t.nr + 1) * kTP + kBand + lane] = f.bot1;
    }
    if (lane < t.nr) {
        tile[(lane + 1) * kTP] = f.left;
        tile[(lane + 1) * kTP + t.nc + 1] = f.right;
    }
}

// Phase 2, step t = 0 .. nr + nc - 2: lane l updates cell (row l, column t - l) of the tile.  The step is a
// dependent chain (north comes from lane l-1's previous step, west from the lane's own), so everything that is
// NOT on the chain is taken off it: the old south / east values and the right-hand side of step t+1 are fetched
// from shared memory during step t (they are overwritten no earlier than step t+2), the lane's previous result
// stays in a register (west) and the north value travels by warp shuffle (the caller passes lane l-1's previous
// result in `north`; lane 0 reads the frame row instead).  The store of the result to shared memory (for the
// write-back and for the next tile's left column) is off the chain as well.
struct StepRegs {
    float west;          // the lane's previous result; initially the frame column left of its row
    float s, e, r, top;  // operands of the coming step: old south, old east, right-hand side, frame row above lane 0
    float* cell;         // the lane's cell of the coming step (unclamped column)
    const float* rp;     // its right-hand side
    const float* tp;     // the frame-row cell above that column
};

F2D_HD void step_fetch(StepRegs& g) {
    g.s = g.cell[kTP];
    g.e = g.cell[1];
    g.top = *g.tp;
    g.r = *g.rp;
}

// `tile` and `rt` must be the two halves of a warp block (kPad floats of slack on either side).
F2D_HD void tile_step_init(const Tile& t, float* tile, const float* rt, int lane, StepRegs& g) {
    const int ll = (lane < t.nr) ? lane : 0;  // rows beyond the band shadow row 0 and never store
    g.west = tile[(ll + 1) * kTP];
    // column q = step - lane, step = 0.  Lanes without a row (lane >= nr) shadow row 0 two columns further left: they
    // then only ever read cells that lanes 0 / 1 wrote in an EARLIER step (a step ends with __syncwarp), never the
    // cell being written in the same step (compute-sanitizer racecheck is clean for bands of one or two rows, too)
    g.cell = tile + (ll + 1) * kTP + (1 - lane) - ((lane < t.nr) ? 0 : 2);
    g.rp = rt + ll * kTileCols - lane;
    g.tp = tile + (1 - lane);
    step_fetch(g);
}

// returns the lane's result of this step (junk if it has no cell this step)
template <bool DIFFUSE>
F2D_HD float tile_step(const Tile& t, int lane, int step, float a, float c, StepRegs& g, float north) {
    const bool on = (lane < t.nr) && ((unsigned)(step - lane) < (unsigned)t.nc);
    float* const here = g.cell;
    const float cs = g.s, ce = g.e, cr = g.r, ct = g.top;
    // operands of step + 1, requested before this step's arithmetic so that their latency hides behind it
    g.cell = here + 1;
    g.rp += 1;
    g.tp += 1;
    step_fetch(g);
    const float n = (lane == 0) ? ct : north;
    const float v = DIFFUSE ? diffuse_cell(n, cs, g.west, ce, cr, a, c) : pressure_cell(n, cs, g.west, ce, cr);
    if (on) {
        *here = v;
        g.west = v;
    }
    return v;
}

F2D_HD float signed_copy(float v, bool negate) { return negate ? -v : v; }

// Phase 3: write the nr x nc results back, plus the edge cells this tile owns.
F2D_HD void tile_store(const Shape& s, const Tile& t, int kind, float* x, const float* tile, int lane) {
    const bool neg_rows = (kind == kBndOppositeVertical);    // top / bottom rows negate (v)
    const bool neg_cols = (kind == kBndOppositeHorizontal);  // left / right columns negate (u)
    float* xr = x + (size_t)t.i0 * s.pitch + t.j0 + lane;
    const float* tc = tile + kTP + lane + 1;
    if (lane < t.nc) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int r = 0; r < kBand; ++r) {
            if (r < t.nr) *xr = tc[r * kTP];
            xr += s.pitch;
        }
    }
    if (lane < t.nc) {
        if (t.i0 == 1) x[t.j0 + lane] = signed_copy(tile[1 * kTP + lane + 1], neg_rows);
        if (t.i0 + t.nr == s.rows - 1)
            x[(size_t)(s.rows - 1) * s.pitch + t.j0 + lane] = signed_copy(tile[t.nr * kTP + lane + 1], neg_rows);
    }
    if (lane < t.nr) {
        float* row = x + (size_t)(t.i0 + lane) * s.pitch;
        if (t.j0 == 1) row[0] = signed_copy(tile[(lane + 1) * kTP + 1], neg_cols);
        if (t.j0 + t.nc == s.cols - 1) row[s.cols - 1] = signed_copy(tile[(lane + 1) * kTP + t.nc], neg_cols);
    }
}

}  // namespace gs
}  // namespace f2d
