// f2d_kernels.cuh -- launcher declarations shared by the step driver (f2d_solver.cu).
#pragma once
#include "f2d_common.cuh"

namespace f2d {

constexpr int kMaxBatch = 3;  // density, u, v relaxed in one launch

// One relaxation problem: next = relax(prev, rhs) with its boundary kind.
struct RelaxField {
    const float* prev;  // iterate k   (nullptr: identically zero -- first pressure pass)
    const float* rhs;   // x0 (diffuse) or divergence (pressure)
    float* next;        // iterate k + sweeps
    float* aux;         // fused first passes, else nullptr.  pressure: prev = u, rhs = v, aux = divergence out;
                        // diffuse: prev = field, rhs = source, aux = x0 out (the field after add_sources)
    int kind;           // F2D_BND_*
    DiffuseCoef coef;   // diffuse only
};
struct RelaxBatch {
    RelaxField f[kMaxBatch];
    int n;
    float dt;  // fused add_sources first diffuse pass (aux != nullptr, prev = field, rhs = source): x0 = FMA(dt, s, f)
};

struct AddSourceBatch {
    const float* f[kMaxBatch];  // field
    float* o[kMaxBatch];        // destination (may alias f: in place)
    const float* s[kMaxBatch];  // source
    int n;
};

// ---- simple one-pass kernels (f2d_kernels_simple.cu) -------------------------------------
void launch_add_sources(const Geom& g, const AddSourceBatch& b, float dt, cudaStream_t st);
void launch_jacobi_naive(const Geom& g, const RelaxBatch& b, bool diffuse, int divmode, cudaStream_t st);
void launch_divergence(const Geom& g, const float* u, const float* v, float* div, float h, cudaStream_t st);
void launch_gradient(const Geom& g, const float* p, const float* u_in, const float* v_in, float* u_out,
                     float* v_out, float h, cudaStream_t st);
// rows [valid_lo, valid_hi) of u0 / v0 are valid; *oob_flag is raised when an owned cell's back-trace leaves them
void launch_advect_velocity(const Geom& g, const float* u0, const float* v0, float* u_out, float* v_out,
                            float dt0, int own_begin, int own_end, int valid_lo, int valid_hi, int* oob_flag, cudaStream_t st);
void launch_add_rows(const Geom& g, float* f, int row0, int nrows, const float* src, cudaStream_t st);
// forward scatter of `src` by (u,v) into `out` (must be zeroed); rows [own_begin, own_end) are the
// source rows this slab owns.  *oob_flag (device int) is set if a target row left the local slab.
void launch_scatter_density(const Geom& g, const float* src, const float* u, const float* v, float* out,
                            float dt0, int own_begin, int own_end, int* oob_flag, cudaStream_t st);
// out = smooth(set_bnd_continuous(in)) on the interior, edges = set_bnd_continuous(in), corners = in
// (do_smooth == false: interior copied; i.e. an out-of-place set_bnd_continuous)
void launch_smooth_bnd(const Geom& g, const float* in, float* out, bool do_smooth, cudaStream_t st);
void launch_smooth_plain(const Geom& g, const float* in, float* out, cudaStream_t st);
void launch_set_bnd_inplace(const Geom& g, float* f, int kind, cudaStream_t st);

// ---- fluid_solver_cpu-compatible stages (F2D_SEM_CPU; f2d_gs.cu) -------------------------------
// One in-place Gauss-Seidel problem: x = relax(x, rhs), `sweeps` lexicographic sweeps with the edge rule of `kind`.
struct GsProblem {
    float* x;          // iterate, updated in place (edges included, corners not)
    const float* rhs;  // x0 (diffuse) or the divergence (pressure); must not alias x
    float a, c;        // diffuse: x = (rhs + a*sum4) / c with c = 1.f + 4.f*a in fp32 (cpp:108); unused for pressure
    int kind;          // F2D_BND_*
};
struct GsBatch {
    GsProblem p[kMaxBatch];
    int n, sweeps;
    int rows, cols, pitch;
    unsigned* flags;   // n * sweeps * bands progress counters, zeroed before the launch ...
    unsigned* ticket;  // ... and the work counter right behind them (gs_flag_words() words in total)
    int* err;          // raised if a wait timed out (reported by f2d_sync)
};
struct CornerBatch {
    float* f[2 * kMaxBatch];
    int n;
};
size_t gs_flag_words(int rows, int nproblems, int sweeps);
void launch_gs_relax(const GsBatch& b, bool diffuse, cudaStream_t st);
void launch_corners_avg(const Geom& g, const CornerBatch& b, cudaStream_t st);
void launch_add_sources_nofma(const Geom& g, const AddSourceBatch& b, float dt, cudaStream_t st);
void launch_advect_velocity_nofma(const Geom& g, const float* u0, const float* v0, float* u_out, float* v_out, float dt0,
                                  cudaStream_t st);
// density scatter in the CPU solver's summation order; writes the interior of `out` only.  keys: one scratch field
// (rows * pitch words), disp_bits: one device word.  Needs rows * pitch < 2^32 - 2.
void launch_scatter_ordered(const Geom& g, const float* src, const float* u, const float* v, float* out, unsigned* keys,
                            float dt0, unsigned* disp_bits, cudaStream_t st);

// ---- peer-to-peer halo exchange (f2d_p2p.cu)
struct XchgSeg {
    const float4* src;  // my rows
    float4* dst;        // the neighbour's rows (peer-mapped)
    unsigned n4;        // float4 count
};
struct XchgParams {
    XchgSeg seg[8];
    int nseg;
    unsigned* my_flags;    // this rank's flag block
    unsigned* up_flags;    // the neighbours' flag blocks (peer-mapped), nullptr at the global edges
    unsigned* down_flags;
    unsigned long long timeout_ns;  // wall-clock bound of every wait
};
void launch_halo_xchg(const XchgParams& p, cudaStream_t st);

// ---- headless renderers (f2d_render.cu)
void launch_density_to_rgba(const Geom& g, const float* d, void* img, float mr, float mg, float mb, cudaStream_t st);
void launch_velocity_to_lines(const Geom& g, const float* u, const float* v, void* lines, float hscale, float vscale,
                              float norm, cudaStream_t st);

// ---- temporally blocked streaming relaxation (f2d_jacobi_stream.cu) ------------------------
struct StreamTuning {
    int chunk_rows;     // output rows per warp (0 = auto)
    int warps_per_cta;  // 0 = auto
    int pdl;            // 1: programmatic dependent launch -- the pass may be scheduled while its predecessor in the stream
                        // drains (it waits with griddepcontrol.wait before it touches global memory)
    int min_blocks;     // unused since round 2
    int min_chunk_mult; // chunks own at least min_chunk_mult * T rows (0 = 2)
    int edge_cost_pct;  // cost of an edge strip's row step in percent of an interior strip's (0 = default)
};
// `sweeps` (1..T) Jacobi sweeps in one pass over the field(s); returns false if (T, geometry)
// is not supported by the streaming kernel.
bool stream_supported(const Geom& g, int T);
cudaError_t launch_jacobi_stream(const Geom& g, const RelaxBatch& b, bool diffuse, int divmode, int T, int sweeps,
                                 const StreamTuning& tune, int sm_count, cudaStream_t st);

}  // namespace f2d
