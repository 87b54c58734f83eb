// gs_emul.cpp -- TEST HARNESS (CPU): runs the tile core of the Gauss-Seidel wavefront kernel
// (fluid-2d_b200/csrc/f2d_gs_tile.h, the very functions the CUDA kernel calls per lane) on the host, one
// "warp" at a time, with the tiles of all sweeps executed in a RANDOM order that respects nothing but the
// kernel's own wait conditions (gs::tile_deps).  If those conditions were insufficient, or a tile touched a
// cell it does not own, some order would read a value of the wrong sweep and the result would differ from
// the oracle's sequential sweep; tests/test_gs_tile_cpu.py checks bit-equality for many seeds.
// This is not a CPU fallback of the product: libf2d.so never contains or calls it.
#include <stdint.h>

#include <vector>

#include "../fluid-2d_b200/csrc/f2d_gs_tile.h"

using namespace f2d::gs;

namespace {
uint64_t next_rand(uint64_t& s) {
    s += 0x9E3779B97F4A7C15ULL;
    uint64_t x = s;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

// on-chip state of one warp = one (sweep, band): shared-memory tiles and the prefetch registers of its 32 lanes
struct Band {
    float block[kWarpFloats];  // the warp block of the kernel: slack, value tile, rhs tile, slack
    float* tile;
    float* rt;
    float pf[32][kPf];
    Band() : tile(block + kPad), rt(block + kPad + kTileFloats) {
        for (int i = 0; i < kWarpFloats; ++i) block[i] = -12345.0f;  // poison: an unloaded cell must never be used
    }
    Band(const Band&) : Band() {}
};

template <bool DIFFUSE>
void run_tile(const Shape& s, Band& b, int w, int c, int kind, float* x, const float* rhs, float a, float cc) {
    Frame fr[32];
    const Tile t = make_tile(s, w, c);
    if (c == 0)  // the first tile has nothing prefetched: fetch now (after the wait)
        for (int lane = 0; lane < 32; ++lane) tile_prefetch(s, t, x, rhs, b.pf[lane], lane);
    for (int lane = 0; lane < 32; ++lane) fr[lane] = tile_frame_load(s, t, x, b.tile, lane, c == 0);
    for (int lane = 0; lane < 32; ++lane) tile_commit(t, b.pf[lane], b.tile, b.rt, lane);
    for (int lane = 0; lane < 32; ++lane) tile_frame_store(t, fr[lane], b.tile, lane);
    if (c + 1 < s.nt) {  // the next tile's interior is read NOW; other tiles run before it is used
        const Tile tn = make_tile(s, w, c + 1);
        for (int lane = 0; lane < 32; ++lane) tile_prefetch(s, tn, x, rhs, b.pf[lane], lane);
    }
    StepRegs g[32];
    float north[32], v[32];
    for (int lane = 0; lane < 32; ++lane) {
        tile_step_init(t, b.tile, b.rt, lane, g[lane]);
        north[lane] = 0.f;
    }
    for (int step = 0; step < t.nr + t.nc - 1; ++step) {
        for (int lane = 0; lane < 32; ++lane) v[lane] = tile_step<DIFFUSE>(t, lane, step, a, cc, g[lane], north[lane]);
        for (int lane = 0; lane < 32; ++lane) north[lane] = lane ? v[lane - 1] : 0.f;  // __shfl_up_sync(v, 1)
    }
    for (int lane = 0; lane < 32; ++lane) tile_store(s, t, kind, x, b.tile, lane);
}
}  // namespace

// K in-place sweeps of x (rows x cols, row pitch `pitch`) with the edge rule of `kind`; corners untouched.
// order: 0 = sequential (k, w, c), otherwise the seed of a random dependency-respecting schedule.
// Returns the number of scheduling attempts that found a band blocked (a measure of how adversarial it was).
extern "C" long gs_emul_relax(float* x, const float* rhs, int rows, int cols, int pitch, int kind, int diffuse,
                              float a, float c, int K, uint64_t order) {
    const Shape s = make_shape(rows, cols, pitch);
    std::vector<unsigned> done((size_t)K * s.nb, 0u);
    std::vector<Band> bands((size_t)K * s.nb);
    long blocked = 0;
    size_t remaining = (size_t)K * s.nb * s.nt;
    uint64_t rng = order;
    size_t cursor = 0;
    while (remaining) {
        // pick a band: sequentially (lowest unfinished) or at random
        size_t b;
        if (order == 0) {
            while (done[cursor] == (unsigned)s.nt) ++cursor;
            b = cursor;
        } else {
            b = (size_t)(next_rand(rng) % done.size());
            if (done[b] == (unsigned)s.nt) continue;
        }
        const int k = (int)(b / s.nb), w = (int)(b % s.nb), ct = (int)done[b];
        const Deps d = tile_deps(s, k, w, ct);
        bool ready = true;
        for (int i = 0; i < d.n; ++i) ready = ready && (done[d.idx[i]] >= d.need[i]);
        if (!ready) {
            ++blocked;
            if (order == 0) return -1;  // the sequential order must never block
            continue;
        }
        if (diffuse)
            run_tile<true>(s, bands[b], w, ct, kind, x, rhs, a, c);
        else
            run_tile<false>(s, bands[b], w, ct, kind, x, rhs, a, c);
        ++done[b];
        --remaining;
    }
    return blocked;
}
