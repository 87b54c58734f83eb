// f2d_solver.cu -- solver object, step driver and the C ABI (include/f2d.h) of libf2d.so.
//
// Replaces the host side of the reference GPU solver: fluid_solver_gpu's constructor
// (src/fluid_solver_gpu.cu:209-218), solve() (:222-258) and the per-stage host methods
// (:260-404), the copy() helpers (src/utilities.hpp:57-83) and kernel_launcher
// (src/kernel_launcher.hpp:8-31).  Differences in mechanism, not in results:
//   * fields live in 128-byte-row-aligned device buffers owned by the handle; ping-pong
//     buffers replace every full-field D2D copy (94 per step in the reference);
//   * boundary passes are fused into the producing kernels (no 1-D launches);
//   * no device-wide synchronisation inside a step (193 per step in the reference): the whole
//     step is captured once into a CUDA graph and replayed;
//   * relaxations run temporally blocked (f2d_jacobi_stream.cu).
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <initializer_list>
#include <new>
#include <utility>
#include <vector>

#include <dlfcn.h>

#include <nvtx3/nvToolsExt.h>  // header-only NVTX v3: ranges cost nothing unless a profiler is attached

#include "f2d_kernels.cuh"

using namespace f2d;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define F2D_CUDA(expr)                                                                              \
    do {                                                                                            \
        cudaError_t e_ = (expr);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            return fail(F2D_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

#define F2D_TRY(expr)            \
    do {                         \
        int rc_ = (expr);        \
        if (rc_ != F2D_OK) return rc_; \
    } while (0)

int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

// NVTX range per stage of the step driver (host-side enqueue / graph-capture time; replays of a captured graph show
// up as one cudaGraphLaunch in a timeline, run with use_graph = 0 to see the stages on the device rows)
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};

struct GraphKey {
    float diffusion_rate, viscosity, dt;
    bool valid;
};

}  // namespace

struct f2d_solver {
    f2d_config cfg;
    Geom g;
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    size_t field_bytes = 0;

    float* state[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // d,u,v,sd,su,sv
    std::vector<float*> temps;                                                   // pool of scratch fields
    std::vector<char> temp_busy;
    float* last_p = nullptr;    // views for F2D_FIELD_PRESSURE / _DIVERGENCE
    float* last_div = nullptr;
    int* oob_flag = nullptr;

    cudaGraphExec_t graph_exec = nullptr;
    GraphKey graph_key = {0.f, 0.f, 0.f, false};
    uint64_t graph_kernels = 0;  // kernel launches inside one replay of the graph
    uint64_t launches = 0;       // kernel launches issued so far (graph nodes included)
    StreamTuning tune = {0, 0, 0, 0, 0, 0};

    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // solve(): the density field is final long before the velocity projections finish; it is copied
    // back on a second stream as soon as ev_density fires
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_density = nullptr, ev_copy = nullptr;
    bool capturing = false;
    bool host_register = false;  // F2D_HOST_REGISTER=1: f2d_solve_host registers every grid it sees (opt-in, see f2d.h)
    void* render_buf = nullptr;  // scratch of the headless renderers
    size_t render_bytes = 0;
    bool fuse_sources = true;     // F2D_FUSE_SOURCES=0: separate add_sources kernel (A/B, cross-check)
    bool fuse_divergence = true;  // F2D_FUSE_DIVERGENCE=0: separate divergence kernel (A/B, cross-check)
    std::vector<std::pair<void*, size_t>> registered;  // host ranges this solver page-locked (f2d_pin_host)
    // solve() pipeline (single GPU): the step cut into four parts, each launched as soon as ITS inputs are uploaded
    //   part 0: add_sources + diffuse u   (needs u, su)      part 2: project, advect, project   (needs parts 0, 1)
    //   part 1: add_sources + diffuse v   (needs v, sv)      part 3: the density chain          (needs d, sd, u, v)
    // The new velocity lands in out_u / out_v (pool buffers held for the life of the solver), so that part 3 still
    // finds the pre-step u, v in the state buffers; it is copied into the state after part 3.
    bool host_pipeline = true;  // F2D_HOST_PIPELINE=0: upload everything, one step graph, download (A/B)
    size_t host_pipeline_min_bytes = 1u << 20;  // smaller fields take the serial order: nothing to hide, 3 extra graph launches (profiles/solve_ab_r01.log)
    cudaStream_t up_stream = nullptr;
    cudaEvent_t ev_in[3] = {nullptr, nullptr, nullptr}, ev_vel = nullptr, ev_fence = nullptr;
    cudaGraphExec_t host_exec[4] = {nullptr, nullptr, nullptr, nullptr};
    GraphKey host_key = {0.f, 0.f, 0.f, false};
    uint64_t host_kernels[4] = {0, 0, 0, 0}, host_xch[4] = {0, 0, 0, 0}, host_xbytes[4] = {0, 0, 0, 0};
    float *out_u = nullptr, *out_v = nullptr;
    float* hp_x0[2] = {nullptr, nullptr};         // add_sources outputs of u, v (alive from part 0/1 to part 2)
    const float* hp_res[2] = {nullptr, nullptr};  // diffused u, v
    // F2D_SEM_CPU (fluid_solver_cpu-compatible arithmetic, f2d_gs.cu)
    unsigned* gs_flags = nullptr;  // progress counters + ticket of the Gauss-Seidel wavefront
    size_t gs_flag_cap = 0;        // words allocated
    unsigned* gs_aux = nullptr;    // [0] wavefront error flag, [1] scatter reach (float bits)
    bool cpu_sem() const { return cfg.semantics == F2D_SEM_CPU; }

    // ---- scratch pool -------------------------------------------------------------------
    float* acquire() {
        for (size_t i = 0; i < temps.size(); ++i)
            if (!temp_busy[i]) {
                temp_busy[i] = 1;
                return temps[i];
            }
        return nullptr;  // sized at create time for the deepest stage; never happens
    }
    void release(const float* p) {
        for (size_t i = 0; i < temps.size(); ++i)
            if (temps[i] == p) temp_busy[i] = 0;
    }

    // ---- scalars of the reference (computed on the host exactly as it does) --------------
    size_t global_cells() const { return (size_t)g.grows * (size_t)g.cols; }
    DiffuseCoef diffuse_coef(float rate, float dt) const {
        DiffuseCoef k;
        float t = dt * (float)global_cells();  // src/fluid_solver_gpu.cu:79, left to right in fp32
        k.a = t * rate;
        k.c = 1.0 + 4.0 * (double)k.a;  // gpu.cu:82
        const double r = 1.0 / k.c;
        k.rc = (float)r;
        k.rl = (float)(r - (double)k.rc);
        return k;
    }
    float dt0(float dt) const { return (float)(std::sqrt((double)global_cells()) * (double)dt); }  // gpu.cu:334
    float h() const { return 1.0f / sqrtf((float)global_cells()); }                                  // gpu.cu:367

    // ---- building blocks -----------------------------------------------------------------
    void count(int n = 1) { launches += (uint64_t)n; }

    // ---- row-slab halo bookkeeping (multi-GPU) -------------------------------------------
    // A slab holds `halo` rows of its neighbours on each slab-internal side.  Every stencil
    // stage is evaluated on ALL local rows; rows closer than the stage radius to a slab-internal
    // edge come out wrong ("invalid").  inv(buf) = number of invalid rows at the slab-internal
    // edges of buf (0 right after an exchange).  A stage with radius r maps inv -> inv + r and is
    // legal while the result stays <= halo (owned rows valid); otherwise its inputs are
    // exchanged first.  One GPU: no neighbours, nothing to do.
    void* comm = nullptr;  // ncclComm_t
    int rank = 0, nranks = 1;
    int cfl_cells = 8;     // bound on the advection displacement per step in rows: the caller's promise (default and
                           // maximum: halo - 1), verified on the device by the advection kernels
    void drop_graphs() {   // the exchange schedule is baked into the captured graphs
        if (graph_exec) cudaGraphExecDestroy(graph_exec);
        graph_exec = nullptr;
        graph_key.valid = false;
        for (auto& e : host_exec) {
            if (e) cudaGraphExecDestroy(e);
            e = nullptr;
        }
        host_key.valid = false;
    }
    float *rx_up = nullptr, *rx_down = nullptr;  // landing zones of the reverse (scatter) exchange
    char* arena = nullptr;                       // the single device allocation all of the above live in
    size_t arena_bytes = 0, field_stride = 0, rx_bytes = 0;
    unsigned* flags = nullptr;                   // P2P transport: epoch + handshake flags (see f2d_p2p.cu)
    struct PeerLink {                            // a neighbour's arena opened through CUDA IPC
        char* arena = nullptr;
        size_t field_stride = 0, rx_bytes = 0;
        int rows = 0, nbuf = 0;
        unsigned* flags() const { return reinterpret_cast<unsigned*>(arena + (size_t)nbuf * field_stride + 2 * rx_bytes); }
        float* rx_up() const { return reinterpret_cast<float*>(arena + (size_t)nbuf * field_stride); }
        float* rx_down() const { return reinterpret_cast<float*>(arena + (size_t)nbuf * field_stride + rx_bytes); }
    };
    PeerLink peer_up, peer_down;
    unsigned long long p2p_timeout_ns = 30ull * 1000000000ull;  // F2D_P2P_TIMEOUT_MS
    bool p2p = false;                            // halo transport: direct peer stores (true) or NCCL (comm != nullptr)
    int nbuffers() const { return 6 + (int)temps.size(); }
    int buffer_index(const float* p) const { return (int)((reinterpret_cast<const char*>(p) - arena) / (ptrdiff_t)field_stride); }
    int exchange_p2p(const float* const* bufs, int n);
    int reverse_exchange_p2p(float* buf);
    uint64_t exchanges = 0, exchanges_in_graph = 0;
    uint64_t xbytes = 0, xbytes_in_graph = 0;  // payload bytes pushed to ONE neighbour (same for up and down)
    std::vector<std::pair<const float*, int>> inv_table;

    bool multi() const { return comm != nullptr || p2p; }
    int H() const { return (int)cfg.halo; }
    bool has_up() const { return g.grow0 > 0; }
    bool has_down() const { return g.grow0 + g.rows < g.grows; }
    int get_inv(const float* p) const {
        for (auto& e : inv_table)
            if (e.first == p) return e.second;
        return 0;
    }
    void set_inv(const float* p, int v) {
        if (!multi() || !p) return;
        for (auto& e : inv_table)
            if (e.first == p) {
                e.second = v;
                return;
            }
        inv_table.emplace_back(p, v);
    }
    int exchange(const float* const* bufs, int n);             // forward halo exchange, inv := 0
    int reverse_exchange_add(float* buf);                      // halo partial sums -> owner, added
    // make sure inv(buf) <= limit for every listed buffer, exchanging the offenders in one group
    int need(std::initializer_list<std::pair<const float*, int>> reqs) {
        if (!multi()) return F2D_OK;
        const float* todo[8];
        int n = 0;
        for (auto& r : reqs) {
            if (!r.first || get_inv(r.first) <= r.second) continue;
            bool dup = false;
            for (int i = 0; i < n; ++i) dup |= (todo[i] == r.first);
            if (!dup && n < 8) todo[n++] = r.first;
        }
        return n ? exchange(todo, n) : F2D_OK;
    }

    // K relaxation sweeps for n problems.  in[i] == nullptr means a zero start (pressure).
    // out[i] receives the buffer holding the final iterate: a pool buffer the caller must
    // release, or in[i] itself when K == 0.
    // With `src` != nullptr (diffuse, stream mode, K > 0) the first pass fuses add_sources: it reads the
    // field in[i] and its source src[i], forms x0 = FMA(dt, s, f), WRITES it to rhs[i] (a pool buffer of
    // the caller) and relaxes from it; the later passes read rhs[i] like any other right-hand side.
    int relax(int n, const float* const* in, const float* const* rhs, const int* kinds, const DiffuseCoef* coefs,
              bool diffuse, uint32_t K, const float** out, const float* const* src = nullptr, float src_dt = 0.f) {
        const float* cur[kMaxBatch];
        float* ping[kMaxBatch] = {nullptr, nullptr, nullptr};
        float* pong[kMaxBatch] = {nullptr, nullptr, nullptr};
        for (int i = 0; i < n; ++i) cur[i] = in[i];
        if (K == 0) {
            for (int i = 0; i < n; ++i) {
                if (cur[i] == nullptr) {  // zero field requested
                    float* z = acquire();
                    if (!z) return fail(F2D_ERR_STATE, "scratch pool exhausted");
                    F2D_CUDA(cudaMemsetAsync(z, 0, field_bytes, stream));
                    cur[i] = z;
                }
                out[i] = cur[i];
            }
            return F2D_OK;
        }
        for (int i = 0; i < n; ++i) {
            ping[i] = acquire();
            pong[i] = acquire();
            if (!ping[i] || !pong[i]) return fail(F2D_ERR_STATE, "scratch pool exhausted");
        }
        const bool stream_mode = (cfg.jacobi_mode == F2D_JACOBI_STREAM);
        uint32_t left = K;
        int flip = 0;
        if (multi()) {
            // the right-hand side is read by every pass: refresh its halos once, up front, together
            // with any start iterate that is not fully valid
            const float* todo[2 * kMaxBatch];
            int m = 0;
            for (int i = 0; i < n; ++i) {
                if (!src && get_inv(rhs[i]) > 0) todo[m++] = rhs[i];
                if (cur[i] && (src || cur[i] != rhs[i]) && get_inv(cur[i]) > 0) todo[m++] = cur[i];
            }
            if (m) F2D_TRY(exchange(todo, m));
        }
        bool fuse_src = (src != nullptr);
        while (left > 0) {
            uint32_t T = 1;
            if (stream_mode) {
                T = diffuse ? cfg.temporal_block_diffuse : cfg.temporal_block;
                while (T > left) T >>= 1;
            }
            if (multi()) {
                // a pass of T sweeps: inv(next) = max(inv(prev) + T, inv(rhs) + T - 1) must stay <= halo
                if ((int)T > H()) return fail(F2D_ERR_INVALID, "halo (%d) shallower than temporal_block (%u)", H(), T);
                const float* todo[2 * kMaxBatch];
                int m = 0;
                for (int i = 0; i < n; ++i) {
                    const bool xp = cur[i] && get_inv(cur[i]) + (int)T > H();
                    const bool xr = !fuse_src && get_inv(rhs[i]) + (int)T - 1 > H();
                    if (xp) todo[m++] = cur[i];
                    if (xr && !(xp && rhs[i] == cur[i])) todo[m++] = rhs[i];
                }
                if (m) F2D_TRY(exchange(todo, m));
            }
            RelaxBatch b;
            b.n = n;
            b.dt = src_dt;
            for (int i = 0; i < n; ++i) {
                b.f[i].prev = cur[i];
                b.f[i].rhs = fuse_src ? src[i] : rhs[i];
                b.f[i].next = flip ? pong[i] : ping[i];
                b.f[i].aux = fuse_src ? const_cast<float*>(rhs[i]) : nullptr;
                b.f[i].kind = kinds[i];
                b.f[i].coef = coefs ? coefs[i] : DiffuseCoef{0.f, 0.f, 0.f, 1.0};
            }
            if (stream_mode) {
                const cudaError_t le = launch_jacobi_stream(g, b, diffuse, (int)cfg.divide_mode, (int)T, (int)T, tune, sm_count, stream);
                if (le != cudaSuccess) return fail(F2D_ERR_CUDA, "streaming Jacobi pass (T=%u) could not be launched: %s", T, cudaGetErrorString(le));
            } else
                launch_jacobi_naive(g, b, diffuse, (int)cfg.divide_mode, stream);
            count();
            for (int i = 0; i < n; ++i) {
                const int ip = cur[i] ? get_inv(cur[i]) : 0;
                if (fuse_src) set_inv(rhs[i], ip);  // x0 is pointwise in the field
                set_inv(b.f[i].next, std::max(ip + (int)T, get_inv(rhs[i]) + (int)T - 1));
                cur[i] = b.f[i].next;
            }
            fuse_src = false;
            flip ^= 1;
            left -= T;
        }
        for (int i = 0; i < n; ++i) {
            out[i] = cur[i];
            release(cur[i] == ping[i] ? pong[i] : ping[i]);
        }
        F2D_CUDA(cudaGetLastError());
        return F2D_OK;
    }

    // project (src/fluid_solver_gpu.cu:358-404): (u_in, v_in) -> (u_out, v_out), all distinct buffers.
    int project(const float* u_in, const float* v_in, float* u_out, float* v_out, uint32_t K) {
        if (last_div) release(last_div);
        if (last_p) release(last_p);
        last_div = last_p = nullptr;
        float* dv = acquire();
        if (!dv) return fail(F2D_ERR_STATE, "scratch pool exhausted");
        const float* pin[1] = {nullptr};
        const float* prhs[1] = {dv};
        const int kind[1] = {F2D_BND_CONTINUOUS};
        const float* pout[1];
        uint32_t k_left = K;
        float* p1 = nullptr;
        if (cfg.jacobi_mode == F2D_JACOBI_STREAM && K > 0 && fuse_divergence) {
            // the divergence is computed inside the first pressure pass (p0 == 0, so sweep 1 is 0.25*div):
            // the pass streams u and v, writes the divergence field for the later passes and p after T sweeps
            uint32_t T = cfg.temporal_block;
            while (T > K) T >>= 1;
            if (multi() && (int)T > H()) return fail(F2D_ERR_INVALID, "halo (%d) shallower than temporal_block (%u)", H(), T);
            F2D_TRY(need({{u_in, H() - (int)T}, {v_in, H() - (int)T}}));
            p1 = acquire();
            if (!p1) return fail(F2D_ERR_STATE, "scratch pool exhausted");
            RelaxBatch b;
            b.n = 1;
            b.dt = 0.f;
            b.f[0].prev = u_in;
            b.f[0].rhs = v_in;
            b.f[0].next = p1;
            b.f[0].aux = dv;
            b.f[0].kind = F2D_BND_CONTINUOUS;
            b.f[0].coef = DiffuseCoef{-0.5f * h(), 0.f, 0.f, 1.0};
            const cudaError_t le = launch_jacobi_stream(g, b, false, (int)cfg.divide_mode, (int)T, (int)T, tune, sm_count, stream);
            if (le != cudaSuccess) return fail(F2D_ERR_CUDA, "fused divergence + pressure pass (T=%u) could not be launched: %s", T, cudaGetErrorString(le));
            count();
            const int iuv = std::max(get_inv(u_in), get_inv(v_in));
            set_inv(dv, iuv + 1);
            set_inv(p1, iuv + (int)T);
            pin[0] = p1;
            k_left = K - T;
        } else {
            F2D_TRY(need({{u_in, H() - 1}, {v_in, H() - 1}}));  // radius-1 stencil
            launch_divergence(g, u_in, v_in, dv, h(), stream);
            count();
            set_inv(dv, std::max(get_inv(u_in), get_inv(v_in)) + 1);
        }
        F2D_TRY(relax(1, pin, prhs, kind, nullptr, false, k_left, pout));
        if (p1 && pout[0] != p1) release(p1);
        F2D_TRY(need({{pout[0], H() - 1}}));
        launch_gradient(g, pout[0], u_in, v_in, u_out, v_out, h(), stream);
        count();
        set_inv(u_out, std::max(get_inv(pout[0]) + 1, std::max(get_inv(u_in), get_inv(v_in))));
        set_inv(v_out, get_inv(u_out));
        F2D_CUDA(cudaGetLastError());
        // keep p / div alive as debug views until the next project
        last_div = dv;
        last_p = const_cast<float*>(pout[0]);
        return F2D_OK;
    }

    // ---- the three chains of a step, shared by the device-resident step and by the pipelined solve() --------------
    // add_sources + diffuse of n fields (gpu.cu:237-238, :243-246).  x0[i]: pool buffer that receives the field after
    // add_sources (== the rhs of the relaxation); with `in_place_first` the first field is updated in place instead
    // (density without the fused first pass).  res[i] = buffer holding the diffused field.
    int diffuse_fields(int n, const int* flds, const int* kinds, const float* rates, float dt, float* const* x0, const float** res) {
        const bool fuse_src = fuse_sources && cfg.jacobi_mode == F2D_JACOBI_STREAM && cfg.diffuse_iters > 0;
        DiffuseCoef kc[kMaxBatch];
        const float* x0c[kMaxBatch];
        for (int i = 0; i < n; ++i) {
            kc[i] = diffuse_coef(rates[i], dt);
            x0c[i] = x0[i];
        }
        if (fuse_src) {
            const float* in[kMaxBatch];
            const float* srcs[kMaxBatch];
            for (int i = 0; i < n; ++i) {
                in[i] = state[flds[i]];
                srcs[i] = state[flds[i] + 3];
            }
            return relax(n, in, x0c, kinds, kc, true, cfg.diffuse_iters, res, srcs, dt);
        }
        AddSourceBatch ab;
        ab.n = n;
        for (int i = 0; i < n; ++i) {
            ab.f[i] = state[flds[i]];
            ab.o[i] = x0[i];
            ab.s[i] = state[flds[i] + 3];
            set_inv(x0[i], get_inv(state[flds[i]]));  // pointwise: halo validity carries over
        }
        launch_add_sources(g, ab, dt, stream);
        count();
        return relax(n, x0c, x0c, kinds, kc, true, cfg.diffuse_iters, res);  // x0 == the field after add_sources
    }

    // density: forward scatter of `dd` (the diffused density) by the PRE-step (u, v), boundary pass + smooth into the
    // density state (gpu.cu:239-240)
    int density_advect(const float* dd, const float* u_pre, const float* v_pre, float dt) {
        float* d = state[F2D_FIELD_DENSITY];
        float* sc = acquire();
        if (!sc) return fail(F2D_ERR_STATE, "scratch pool exhausted");
        F2D_CUDA(cudaMemsetAsync(sc, 0, field_bytes, stream));  // gpu.cu:337
        launch_scatter_density(g, dd, u_pre, v_pre, sc, dt0(dt), own_begin(), own_end(), oob_flag, stream);
        count();
        if (multi()) {
            // splats that landed in halo rows belong to the neighbour slab: send them home and add,
            // then refresh the halos for the radius-1 smooth
            F2D_TRY(reverse_exchange_add(sc));
            const float* one[1] = {sc};
            F2D_TRY(exchange(one, 1));
        }
        launch_smooth_bnd(g, sc, d, cfg.smooth != 0, stream);  // gpu.cu:355 + :240
        count();
        set_inv(d, 1);
        release(sc);
        F2D_CUDA(cudaGetLastError());
        return F2D_OK;
    }

    // velocity: project, self-advect, project (gpu.cu:247-252).  (u1, v1): the diffused velocity in pool buffers, which
    // are overwritten by the advection; the result lands in (u_out, v_out).
    int velocity_chain(const float* u1, const float* v1, float* u_out, float* v_out, float dt) {
        float *u2 = acquire(), *v2 = acquire();
        if (!u2 || !v2) return fail(F2D_ERR_STATE, "scratch pool exhausted");
        F2D_TRY(project(u1, v1, u2, v2, cfg.project_iters));  // gpu.cu:247
        // advect both components by (U0,V0) = (u2,v2) (gpu.cu:248-251); u1/v1 are free to be overwritten
        float *u3 = const_cast<float*>(u1), *v3 = const_cast<float*>(v1);
        {
            // gather radius = displacement bound + bilinear footprint.  The bound is the caller's (cfl_cells, at most
            // halo - 1) and is VERIFIED on the device: a back-trace that leaves the rows known to be valid raises the
            // error flag f2d_sync reports, it never reads stale halo rows silently.
            const int r = multi() ? cfl_cells + 1 : 0;
            F2D_TRY(need({{u2, H() - r}, {v2, H() - r}}));
            const int iuv = std::max(get_inv(u2), get_inv(v2));
            const int valid_lo = has_up() ? iuv : 0, valid_hi = has_down() ? g.rows - iuv : g.rows;
            launch_advect_velocity(g, u2, v2, u3, v3, dt0(dt), own_begin(), own_end(), valid_lo, valid_hi, oob_flag, stream);
            count();
            set_inv(u3, iuv + r);
            set_inv(v3, iuv + r);
        }
        release(u2);
        release(v2);
        F2D_TRY(project(u3, v3, u_out, v_out, cfg.project_iters));  // gpu.cu:252
        F2D_CUDA(cudaGetLastError());
        return F2D_OK;
    }

    // halos of the state are assumed stale at the start of a step (only owned rows valid), so that the exchange
    // schedule baked into a captured graph is right for every replay
    int begin_step_bookkeeping() {
        if (!multi()) return F2D_OK;
        inv_table.clear();
        set_inv(state[F2D_FIELD_DENSITY], H());
        set_inv(state[F2D_FIELD_U], H());
        set_inv(state[F2D_FIELD_V], H());
        if (cfl_cells + 1 > H()) return fail(F2D_ERR_INVALID, "halo (%d) shallower than the advection radius (%d)", H(), cfl_cells + 1);
        return F2D_OK;
    }

    int record_density_done() {
        if (capturing)
            F2D_CUDA(cudaEventRecordWithFlags(ev_density, stream, cudaEventRecordExternal));
        else
            F2D_CUDA(cudaEventRecord(ev_density, stream));
        return F2D_OK;
    }

    // One full solve() step on the device-resident state (order of gpu.cu:236-252).
    int enqueue_step(float diffusion_rate, float viscosity, float dt) {
        float *d = state[F2D_FIELD_DENSITY], *u = state[F2D_FIELD_U], *v = state[F2D_FIELD_V];
        F2D_TRY(begin_step_bookkeeping());
        NvtxRange step_range("f2d step");
        // ---- add_sources + diffuse for d, u, v in one batch.  u and v get their sources OUT of place: the density
        //      scatter below still needs the pre-step u, v (the reference runs the whole density chain first,
        //      gpu.cu:236-240).  In stream mode add_sources is fused into the first diffuse pass, which writes x0.
        const bool fuse_src = fuse_sources && cfg.jacobi_mode == F2D_JACOBI_STREAM && cfg.diffuse_iters > 0;
        float *ds = fuse_src ? acquire() : d, *us = acquire(), *vs = acquire();
        if (!ds || !us || !vs) return fail(F2D_ERR_STATE, "scratch pool exhausted");
        const float* dif[3];
        {
            NvtxRange r("add_sources + diffuse (d, u, v)");
            const int flds[3] = {F2D_FIELD_DENSITY, F2D_FIELD_U, F2D_FIELD_V};
            const int kind[3] = {F2D_BND_CONTINUOUS, F2D_BND_OPPOSITE_HORIZONTAL, F2D_BND_OPPOSITE_VERTICAL};
            const float rates[3] = {diffusion_rate, viscosity, viscosity};
            float* x0[3] = {ds, us, vs};
            F2D_TRY(diffuse_fields(3, flds, kind, rates, dt, x0, dif));
        }
        {
            NvtxRange r("density: scatter + smooth");
            F2D_TRY(density_advect(dif[0], u, v, dt));
            if (dif[0] != d) release(dif[0]);
            if (ds != d && ds != dif[0]) release(ds);
            // density is final: let solve() start its download while the projections run
            F2D_TRY(record_density_done());
        }
        {
            NvtxRange r("velocity: project, advect, project");
            const float *u1 = dif[1], *v1 = dif[2];
            if (u1 != us) {  // K > 0: the add_sources outputs are no longer needed
                release(us);
                release(vs);
            }
            F2D_TRY(velocity_chain(u1, v1, u, v, dt));  // result lands in the state buffers
            release(u1);
            release(v1);
        }
        F2D_CUDA(cudaGetLastError());
        return F2D_OK;
    }

    // ================================================================ F2D_SEM_CPU: fluid_solver_cpu, bit for bit
    float dt0_cpu(float dt) const { return sqrtf((float)global_cells()) * dt; }  // cpp:126 (float sqrt, float product)

    int ensure_gs_flags(size_t words) {
        if (words <= gs_flag_cap) return F2D_OK;
        if (capturing) return fail(F2D_ERR_STATE, "Gauss-Seidel flag block too small inside a graph capture");
        F2D_CUDA(cudaStreamSynchronize(stream));
        if (gs_flags) cudaFree(gs_flags);
        gs_flags = nullptr;
        gs_flag_cap = 0;
        F2D_CUDA(cudaMalloc(&gs_flags, words * sizeof(unsigned)));
        gs_flag_cap = words;
        return F2D_OK;
    }

    int corners(std::initializer_list<float*> fields) {
        CornerBatch cb;
        cb.n = 0;
        for (float* f : fields) cb.f[cb.n++] = f;
        launch_corners_avg(g, cb, stream);
        count();
        F2D_CUDA(cudaGetLastError());
        return F2D_OK;
    }

    // K in-place lexicographic Gauss-Seidel sweeps of n fields (one wavefront launch), each followed by the
    // boundary pass of its kind (cpp:104-113 / :196-204); the corners are averaged once at the end, which is
    // what K boundary passes leave behind (no stencil reads a corner in between).
    int relax_gs(int n, float* const* x, const float* const* rhs, const int* kinds, const float* a, bool diffuse, uint32_t K) {
        if (K == 0) return F2D_OK;
        const size_t words = gs_flag_words(g.rows, n, (int)K);
        F2D_TRY(ensure_gs_flags(words));
        F2D_CUDA(cudaMemsetAsync(gs_flags, 0, words * sizeof(unsigned), stream));
        GsBatch b;
        b.n = n;
        b.sweeps = (int)K;
        b.rows = g.rows;
        b.cols = g.cols;
        b.pitch = g.pitch;
        b.flags = gs_flags;
        b.ticket = gs_flags + (words - 1);
        b.err = reinterpret_cast<int*>(gs_aux);
        CornerBatch cb;
        cb.n = n;
        for (int i = 0; i < n; ++i) {
            b.p[i].x = x[i];
            b.p[i].rhs = rhs[i];
            b.p[i].a = a ? a[i] : 0.f;
            b.p[i].c = a ? 1.f + 4.f * a[i] : 1.f;  // cpp:108, fp32
            b.p[i].kind = kinds[i];
            cb.f[i] = x[i];
        }
        launch_gs_relax(b, diffuse, stream);
        count();
        launch_corners_avg(g, cb, stream);
        count();
        F2D_CUDA(cudaGetLastError());
        return F2D_OK;
    }

    // diffuse (cpp:95-114) of n fields in place: x0 <- field, K sweeps
    int diffuse_cpu(int n, float* const* x, const int* kinds, const float* rates, float dt, uint32_t K) {
        if (K == 0) return F2D_OK;
        float* x0[kMaxBatch] = {nullptr, nullptr, nullptr};
        const float* rhs[kMaxBatch];
        float a[kMaxBatch];
        for (int i = 0; i < n; ++i) {
            x0[i] = acquire();
            if (!x0[i]) return fail(F2D_ERR_STATE, "scratch pool exhausted");
            F2D_CUDA(cudaMemcpyAsync(x0[i], x[i], field_bytes, cudaMemcpyDeviceToDevice, stream));  // cpp:102
            rhs[i] = x0[i];
            a[i] = diffuse_coef(rates[i], dt).a;  // cpp:100, the same fp32 expression as gpu.cu:79
        }
        int rc = relax_gs(n, x, rhs, kinds, a, true, K);
        for (int i = 0; i < n; ++i) release(x0[i]);
        return rc;
    }

    // project (cpp:179-215): (u_in, v_in) -> (u_out, v_out), all distinct buffers
    int project_cpu(const float* u_in, const float* v_in, float* u_out, float* v_out, uint32_t K) {
        if (last_div) release(last_div);
        if (last_p) release(last_p);
        last_div = last_p = nullptr;
        float *dv = acquire(), *p = acquire();
        if (!dv || !p) return fail(F2D_ERR_STATE, "scratch pool exhausted");
        launch_divergence(g, u_in, v_in, dv, h(), stream);  // cpp:190-193 + edges of :195
        count();
        F2D_TRY(corners({dv}));
        F2D_CUDA(cudaMemsetAsync(p, 0, field_bytes, stream));  // cpp:186
        float* px[1] = {p};
        const float* prhs[1] = {dv};
        const int kind[1] = {F2D_BND_CONTINUOUS};
        F2D_TRY(relax_gs(1, px, prhs, kind, nullptr, false, K));
        launch_gradient(g, p, u_in, v_in, u_out, v_out, h(), stream);  // cpp:207-211 + edges of :213-214
        count();
        F2D_TRY(corners({u_out, v_out}));
        last_div = dv;
        last_p = p;
        return F2D_OK;
    }

    // density advect (cpp:127-152 + :176): f <- scatter of f by (u, v) with the boundary pass.  `do_smooth` adds
    // fluid_solver_gpu's density smooth (gpu.cu:314-323); fluid_solver_cpu has none.
    int advect_density_cpu(float* f, const float* u, const float* v, float dt, bool do_smooth) {
        float *sc = acquire(), *keys = acquire();
        if (!sc || !keys) return fail(F2D_ERR_STATE, "scratch pool exhausted");
        launch_scatter_ordered(g, f, u, v, sc, reinterpret_cast<unsigned*>(keys), dt0_cpu(dt), gs_aux + 1, stream);  // interior of sc
        count(2);
        release(keys);
        launch_smooth_bnd(g, sc, f, do_smooth, stream);  // edges from the scattered interior, out of place
        count();
        release(sc);
        return corners({f});
    }

    // One full fluid_solver_cpu::solve step (cpp:15-30) on the device-resident state.
    int enqueue_step_cpu(float diffusion_rate, float viscosity, float dt) {
        float *d = state[F2D_FIELD_DENSITY], *u = state[F2D_FIELD_U], *v = state[F2D_FIELD_V];
        float *us = acquire(), *vs = acquire();
        if (!us || !vs) return fail(F2D_ERR_STATE, "scratch pool exhausted");
        // add_sources (cpp:16, :21-22).  u and v go out of place: the density advect below needs the pre-step u, v
        AddSourceBatch ab;
        ab.n = 3;
        ab.f[0] = d;
        ab.o[0] = d;
        ab.s[0] = state[F2D_FIELD_DENSITY_SOURCE];
        ab.f[1] = u;
        ab.o[1] = us;
        ab.s[1] = state[F2D_FIELD_U_SOURCE];
        ab.f[2] = v;
        ab.o[2] = vs;
        ab.s[2] = state[F2D_FIELD_V_SOURCE];
        launch_add_sources_nofma(g, ab, dt, stream);
        count();
        // the three diffuses (cpp:17, :23-24) share one wavefront launch
        {
            float* x[3] = {d, us, vs};
            const int kind[3] = {F2D_BND_CONTINUOUS, F2D_BND_OPPOSITE_HORIZONTAL, F2D_BND_OPPOSITE_VERTICAL};
            const float rates[3] = {diffusion_rate, viscosity, viscosity};
            F2D_TRY(diffuse_cpu(3, x, kind, rates, dt, cfg.diffuse_iters));
        }
        // density advect by the pre-step velocity (cpp:18)
        F2D_TRY(advect_density_cpu(d, u, v, dt, cfg.smooth != 0));
        if (capturing)
            F2D_CUDA(cudaEventRecordWithFlags(ev_density, stream, cudaEventRecordExternal));
        else
            F2D_CUDA(cudaEventRecord(ev_density, stream));
        // velocity: project, self-advect, project (cpp:25-30)
        float *u2 = acquire(), *v2 = acquire();
        if (!u2 || !v2) return fail(F2D_ERR_STATE, "scratch pool exhausted");
        F2D_TRY(project_cpu(us, vs, u2, v2, cfg.project_iters));
        launch_advect_velocity_nofma(g, u2, v2, us, vs, dt0_cpu(dt), stream);  // cpp:26-29
        count();
        F2D_TRY(corners({us, vs}));
        release(u2);
        release(v2);
        F2D_TRY(project_cpu(us, vs, u, v, cfg.project_iters));
        release(us);
        release(vs);
        F2D_CUDA(cudaGetLastError());
        return F2D_OK;
    }

    // ================================================================ solve() pipeline parts (F2D_SEM_GPU; one GPU or one slab)
    int enqueue_host_part(int part, float diffusion_rate, float viscosity, float dt) {
        float *d = state[F2D_FIELD_DENSITY], *u = state[F2D_FIELD_U], *v = state[F2D_FIELD_V];
        const bool fuse_src = fuse_sources && cfg.jacobi_mode == F2D_JACOBI_STREAM && cfg.diffuse_iters > 0;
        if (part == 0) F2D_TRY(begin_step_bookkeeping());  // the four parts are always enqueued in order 0, 1, 2, 3
        if (part == 0 || part == 1) {  // gpu.cu:243-246 for one velocity component
            NvtxRange r(part == 0 ? "solve part 0: add_sources + diffuse u" : "solve part 1: add_sources + diffuse v");
            const int flds[1] = {part == 0 ? F2D_FIELD_U : F2D_FIELD_V};
            const int kind[1] = {part == 0 ? F2D_BND_OPPOSITE_HORIZONTAL : F2D_BND_OPPOSITE_VERTICAL};
            const float rates[1] = {viscosity};
            float* x0[1] = {acquire()};
            if (!x0[0]) return fail(F2D_ERR_STATE, "scratch pool exhausted");
            const float* res[1];
            F2D_TRY(diffuse_fields(1, flds, kind, rates, dt, x0, res));
            hp_x0[part] = x0[0];
            hp_res[part] = res[0];
            return F2D_OK;
        }
        if (part == 2) {  // gpu.cu:247-252
            NvtxRange r("solve part 2: project, advect, project");
            const float *u1 = hp_res[0], *v1 = hp_res[1];
            if (u1 != hp_x0[0]) release(hp_x0[0]);
            if (v1 != hp_x0[1]) release(hp_x0[1]);
            F2D_TRY(velocity_chain(u1, v1, out_u, out_v, dt));
            release(u1);
            release(v1);
            hp_x0[0] = hp_x0[1] = nullptr;
            hp_res[0] = hp_res[1] = nullptr;
            return F2D_OK;
        }
        // part 3: add_sources + diffuse + scatter + smooth of the density (gpu.cu:237-240), by the PRE-step u, v
        NvtxRange r("solve part 3: density chain");
        const int flds[1] = {F2D_FIELD_DENSITY};
        const int kind[1] = {F2D_BND_CONTINUOUS};
        const float rates[1] = {diffusion_rate};
        float* x0[1] = {fuse_src ? acquire() : d};
        if (!x0[0]) return fail(F2D_ERR_STATE, "scratch pool exhausted");
        const float* dif[1];
        F2D_TRY(diffuse_fields(1, flds, kind, rates, dt, x0, dif));
        F2D_TRY(density_advect(dif[0], u, v, dt));
        if (dif[0] != d) release(dif[0]);
        if (x0[0] != d && x0[0] != dif[0]) release(x0[0]);
        return F2D_OK;
    }

    int ensure_host_graphs(float diffusion_rate, float viscosity, float dt) {
        if (host_exec[0] && host_key.valid && host_key.diffusion_rate == diffusion_rate && host_key.viscosity == viscosity &&
            host_key.dt == dt)
            return F2D_OK;
        for (auto& e : host_exec) {
            if (e) cudaGraphExecDestroy(e);
            e = nullptr;
        }
        host_key.valid = false;
        if (last_div) release(last_div);
        if (last_p) release(last_p);
        last_div = last_p = nullptr;
        for (int part = 0; part < 4; ++part) {
            const uint64_t before = launches, xbefore = exchanges, bbefore = xbytes;
            cudaGraph_t graph = nullptr;
            // relaxed mode: NCCL (multi-GPU) may issue its own runtime calls while we capture
            F2D_CUDA(cudaStreamBeginCapture(stream, multi() ? cudaStreamCaptureModeRelaxed : cudaStreamCaptureModeThreadLocal));
            capturing = true;
            int rc = enqueue_host_part(part, diffusion_rate, viscosity, dt);
            capturing = false;
            cudaError_t ce = cudaStreamEndCapture(stream, &graph);
            host_kernels[part] = launches - before;
            launches = before;
            host_xch[part] = exchanges - xbefore;
            exchanges = xbefore;
            host_xbytes[part] = xbytes - bbefore;
            xbytes = bbefore;
            if (rc != F2D_OK) {
                if (graph) cudaGraphDestroy(graph);
                return rc;
            }
            if (ce != cudaSuccess) return fail(F2D_ERR_CUDA, "cudaStreamEndCapture failed: %s", cudaGetErrorString(ce));
            ce = cudaGraphInstantiate(&host_exec[part], graph, 0);
            cudaGraphDestroy(graph);
            if (ce != cudaSuccess) return fail(F2D_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ce));
        }
        host_key = {diffusion_rate, viscosity, dt, true};
        return F2D_OK;
    }

    int run_host_part(int part, float diffusion_rate, float viscosity, float dt) {
        if (!cfg.use_graph) return enqueue_host_part(part, diffusion_rate, viscosity, dt);
        F2D_CUDA(cudaGraphLaunch(host_exec[part], stream));
        launches += host_kernels[part];
        exchanges += host_xch[part];
        xbytes += host_xbytes[part];
        return F2D_OK;
    }

    int h2d_on(float* dst, const float* src, cudaStream_t st) {
        return cudaMemcpy2DAsync(dst, (size_t)g.pitch * sizeof(float), src, (size_t)g.cols * sizeof(float),
                                 (size_t)g.cols * sizeof(float), (size_t)g.rows, cudaMemcpyHostToDevice, st) == cudaSuccess
                   ? F2D_OK
                   : fail(F2D_ERR_CUDA, "cudaMemcpy2DAsync(H2D) failed: %s", cudaGetErrorString(cudaGetLastError()));
    }

    // fluid_solver::solve with the uploads, the four parts of the step and the downloads overlapped: the copy
    // engines run in both directions while the SMs work on whatever already arrived.
    int solve_host_pipelined(float* density, const float* density_source, float diffusion_rate, float* u, float* v,
                             const float* u_source, const float* v_source, float viscosity, float dt) {
        if (!out_u) {
            out_u = acquire();
            out_v = acquire();
            if (!out_u || !out_v) return fail(F2D_ERR_STATE, "scratch pool exhausted");
        }
        if (cfg.use_graph) F2D_TRY(ensure_host_graphs(diffusion_rate, viscosity, dt));
        // uploads, in the order the parts need them (whatever is queued on the solver stream goes first)
        F2D_CUDA(cudaEventRecord(ev_fence, stream));
        F2D_CUDA(cudaStreamWaitEvent(up_stream, ev_fence, 0));
        F2D_TRY(h2d_on(state[F2D_FIELD_U], u, up_stream));
        F2D_TRY(h2d_on(state[F2D_FIELD_U_SOURCE], u_source, up_stream));
        F2D_CUDA(cudaEventRecord(ev_in[0], up_stream));
        F2D_TRY(h2d_on(state[F2D_FIELD_V], v, up_stream));
        F2D_TRY(h2d_on(state[F2D_FIELD_V_SOURCE], v_source, up_stream));
        F2D_CUDA(cudaEventRecord(ev_in[1], up_stream));
        F2D_TRY(h2d_on(state[F2D_FIELD_DENSITY_SOURCE], density_source, up_stream));
        F2D_TRY(h2d_on(state[F2D_FIELD_DENSITY], density, up_stream));
        F2D_CUDA(cudaEventRecord(ev_in[2], up_stream));
        // compute
        F2D_CUDA(cudaStreamWaitEvent(stream, ev_in[0], 0));
        F2D_TRY(run_host_part(0, diffusion_rate, viscosity, dt));
        F2D_CUDA(cudaStreamWaitEvent(stream, ev_in[1], 0));
        F2D_TRY(run_host_part(1, diffusion_rate, viscosity, dt));
        F2D_TRY(run_host_part(2, diffusion_rate, viscosity, dt));
        F2D_CUDA(cudaEventRecord(ev_vel, stream));
        F2D_CUDA(cudaStreamWaitEvent(stream, ev_in[2], 0));
        F2D_TRY(run_host_part(3, diffusion_rate, viscosity, dt));
        F2D_CUDA(cudaEventRecord(ev_density, stream));
        // the device-resident state follows the host grids: u, v <- the new velocity
        F2D_CUDA(cudaMemcpyAsync(state[F2D_FIELD_U], out_u, field_bytes, cudaMemcpyDeviceToDevice, stream));
        F2D_CUDA(cudaMemcpyAsync(state[F2D_FIELD_V], out_v, field_bytes, cudaMemcpyDeviceToDevice, stream));
        // downloads: velocity as soon as part 2 is done (the density chain is still running), then the density
        F2D_CUDA(cudaStreamWaitEvent(copy_stream, ev_vel, 0));
        F2D_TRY(d2h_on(u, out_u, copy_stream));
        F2D_TRY(d2h_on(v, out_v, copy_stream));
        F2D_CUDA(cudaStreamWaitEvent(copy_stream, ev_density, 0));
        F2D_TRY(d2h_on(density, state[F2D_FIELD_DENSITY], copy_stream));
        F2D_CUDA(cudaEventRecord(ev_copy, copy_stream));
        F2D_CUDA(cudaStreamWaitEvent(stream, ev_copy, 0));  // nothing queued later may overtake the downloads
        F2D_CUDA(cudaStreamSynchronize(copy_stream));
        return F2D_OK;
    }

    // Size that decides between the pipelined and the serial solve().  The two issue DIFFERENT halo-exchange schedules
    // (fields one by one vs batched), so every rank of a slab run must take the same path: the decision uses the
    // nominal interior slab (rows / ranks + two halos), not this rank's own size -- edge slabs are one halo smaller.
    size_t pipeline_decision_bytes() const {
        if (!multi()) return field_bytes;
        const size_t nominal_rows = ((size_t)g.grows + (size_t)nranks - 1) / (size_t)nranks + 2 * (size_t)H();
        return nominal_rows * (size_t)g.pitch * sizeof(float);
    }

    // rows of the local slab whose cells this solver owns (halo rows excluded)
    int own_begin() const { return (g.grow0 == 0) ? 0 : (int)cfg.halo; }
    int own_end() const { return (g.grow0 + g.rows == g.grows) ? g.rows : g.rows - (int)cfg.halo; }

    int ensure_graph(float diffusion_rate, float viscosity, float dt) {
        if (graph_exec && graph_key.valid && graph_key.diffusion_rate == diffusion_rate &&
            graph_key.viscosity == viscosity && graph_key.dt == dt)
            return F2D_OK;
        if (graph_exec) {
            cudaGraphExecDestroy(graph_exec);
            graph_exec = nullptr;
        }
        // views from an earlier eager stage would otherwise pin pool buffers forever
        if (last_div) release(last_div);
        if (last_p) release(last_p);
        last_div = last_p = nullptr;
        const uint64_t before = launches, xbefore = exchanges, bbefore = xbytes;
        cudaGraph_t graph = nullptr;
        // relaxed mode: NCCL (multi-GPU) may issue its own runtime calls while we capture
        F2D_CUDA(cudaStreamBeginCapture(stream, multi() ? cudaStreamCaptureModeRelaxed : cudaStreamCaptureModeThreadLocal));
        capturing = true;
        int rc = cpu_sem() ? enqueue_step_cpu(diffusion_rate, viscosity, dt) : enqueue_step(diffusion_rate, viscosity, dt);
        capturing = false;
        cudaError_t ce = cudaStreamEndCapture(stream, &graph);
        if (rc != F2D_OK) {
            if (graph) cudaGraphDestroy(graph);
            return rc;
        }
        if (ce != cudaSuccess) return fail(F2D_ERR_CUDA, "cudaStreamEndCapture failed: %s", cudaGetErrorString(ce));
        graph_kernels = launches - before;
        launches = before;  // capture launched nothing yet
        exchanges_in_graph = exchanges - xbefore;
        exchanges = xbefore;
        xbytes_in_graph = xbytes - bbefore;
        xbytes = bbefore;
        F2D_CUDA(cudaGraphInstantiate(&graph_exec, graph, 0));
        cudaGraphDestroy(graph);
        graph_key = {diffusion_rate, viscosity, dt, true};
        return F2D_OK;
    }

    int step(float diffusion_rate, float viscosity, float dt, uint32_t nsteps) {
        F2D_CUDA(cudaSetDevice(device));
        if (cfg.use_graph) {
            F2D_TRY(ensure_graph(diffusion_rate, viscosity, dt));
            for (uint32_t s = 0; s < nsteps; ++s) {
                F2D_CUDA(cudaGraphLaunch(graph_exec, stream));
                launches += graph_kernels;
                exchanges += exchanges_in_graph;
                xbytes += xbytes_in_graph;
            }
        } else {
            for (uint32_t s = 0; s < nsteps; ++s)
                F2D_TRY(cpu_sem() ? enqueue_step_cpu(diffusion_rate, viscosity, dt) : enqueue_step(diffusion_rate, viscosity, dt));
        }
        return F2D_OK;
    }

    // F2D_HOST_REGISTER=1 only (the caller then promises that every grid outlives the solver, f2d.h): page-lock a
    // host grid on first sight.  The explicit, scoped way is f2d_pin_host / f2d_unpin_host.  Best effort.
    void pin_host(const void* p, size_t bytes) {
        if (!host_register || !p || bytes < (1u << 20)) return;  // small grids: the staged copy is cheap
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
            cudaGetLastError();
            return;
        }
        if (at.type != cudaMemoryTypeUnregistered) return;
        if (cudaHostRegister(const_cast<void*>(p), bytes, cudaHostRegisterDefault) == cudaSuccess)
            registered.emplace_back(const_cast<void*>(p), bytes);
        else
            cudaGetLastError();
    }

    int d2h_on(float* dst, const float* src, cudaStream_t st) {
        return cudaMemcpy2DAsync(dst, (size_t)g.cols * sizeof(float), src, (size_t)g.pitch * sizeof(float),
                                 (size_t)g.cols * sizeof(float), (size_t)g.rows, cudaMemcpyDeviceToHost, st) == cudaSuccess
                   ? F2D_OK
                   : fail(F2D_ERR_CUDA, "cudaMemcpy2DAsync(D2H) failed: %s", cudaGetErrorString(cudaGetLastError()));
    }

    int h2d(float* dst, const float* src) {
        return cudaMemcpy2DAsync(dst, (size_t)g.pitch * sizeof(float), src, (size_t)g.cols * sizeof(float),
                                 (size_t)g.cols * sizeof(float), (size_t)g.rows, cudaMemcpyHostToDevice, stream) == cudaSuccess
                   ? F2D_OK
                   : fail(F2D_ERR_CUDA, "cudaMemcpy2DAsync(H2D) failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    int d2h(float* dst, const float* src) {
        return cudaMemcpy2DAsync(dst, (size_t)g.cols * sizeof(float), src, (size_t)g.pitch * sizeof(float),
                                 (size_t)g.cols * sizeof(float), (size_t)g.rows, cudaMemcpyDeviceToHost, stream) == cudaSuccess
                   ? F2D_OK
                   : fail(F2D_ERR_CUDA, "cudaMemcpy2DAsync(D2H) failed: %s", cudaGetErrorString(cudaGetLastError()));
    }

    float* field_ptr(int field) {
        if (field >= 0 && field < 6) return state[field];
        if (field == F2D_FIELD_PRESSURE) return last_p;
        if (field == F2D_FIELD_DIVERGENCE) return last_div;
        return nullptr;
    }
};

// ============================================================ halo exchange over NCCL (NVLink)
// NCCL is resolved at run time (dlopen) so that libf2d.so has no link-time dependency on it and a
// process that already loaded a copy (torch bundles one) shares it.  The send/recv pairs are
// enqueued on the solver's stream, so they are captured into the step's CUDA graph like kernels.
namespace {
struct NcclId {
    char internal[128];
};
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(NcclId*) = nullptr;
    int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
constexpr int kNcclFloat = 7;  // ncclFloat32 (nccl.h)

int load_nccl() {
    if (g_nccl.lib) return F2D_OK;
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);  // a copy already in the process
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return fail(F2D_ERR_STATE, "cannot load libnccl.so.2: %s", dlerror());
    auto sym = [&](const char* n) { return dlsym(lib, n); };
    g_nccl.GetUniqueId = (int (*)(NcclId*))sym("ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(void**, int, NcclId, int))sym("ncclCommInitRank");
    g_nccl.CommDestroy = (int (*)(void*))sym("ncclCommDestroy");
    g_nccl.Send = (int (*)(const void*, size_t, int, int, void*, cudaStream_t))sym("ncclSend");
    g_nccl.Recv = (int (*)(void*, size_t, int, int, void*, cudaStream_t))sym("ncclRecv");
    g_nccl.GroupStart = (int (*)())sym("ncclGroupStart");
    g_nccl.GroupEnd = (int (*)())sym("ncclGroupEnd");
    g_nccl.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.Send || !g_nccl.Recv || !g_nccl.GroupStart ||
        !g_nccl.GroupEnd || !g_nccl.GetErrorString || !g_nccl.CommDestroy)
        return fail(F2D_ERR_STATE, "libnccl.so.2 lacks a required symbol");
    g_nccl.lib = lib;
    return F2D_OK;
}

#define F2D_NCCL(expr)                                                                                   \
    do {                                                                                                 \
        int r_ = (expr);                                                                                 \
        if (r_ != 0) return fail(F2D_ERR_CUDA, "%s failed: %s", #expr, g_nccl.GetErrorString(r_));      \
    } while (0)
}  // namespace

// Forward exchange: my first/last `halo` OWNED rows go to the neighbours' halo rows, theirs come
// into mine.  All listed buffers travel in one NCCL group (one fused P2P kernel over NVLink).
int f2d_solver::exchange(const float* const* bufs, int n) {
    if (!multi() || n == 0) return F2D_OK;
    if (p2p) return exchange_p2p(bufs, n);
    const size_t cnt = (size_t)H() * (size_t)g.pitch;
    F2D_NCCL(g_nccl.GroupStart());
    for (int i = 0; i < n; ++i) {
        float* b = const_cast<float*>(bufs[i]);
        if (has_up()) {
            F2D_NCCL(g_nccl.Send(b + cnt, cnt, kNcclFloat, rank - 1, comm, stream));  // rows [H, 2H)
            F2D_NCCL(g_nccl.Recv(b, cnt, kNcclFloat, rank - 1, comm, stream));        // rows [0, H)
        }
        if (has_down()) {
            float* tail = b + (size_t)(g.rows - 2 * H()) * g.pitch;
            F2D_NCCL(g_nccl.Send(tail, cnt, kNcclFloat, rank + 1, comm, stream));        // rows [rows-2H, rows-H)
            F2D_NCCL(g_nccl.Recv(tail + cnt, cnt, kNcclFloat, rank + 1, comm, stream));  // rows [rows-H, rows)
        }
    }
    F2D_NCCL(g_nccl.GroupEnd());
    for (int i = 0; i < n; ++i) set_inv(bufs[i], 0);
    ++exchanges;
    xbytes += (uint64_t)n * cnt * sizeof(float);
    return F2D_OK;
}

// Reverse exchange of the density scatter: the partial sums that landed in my halo rows belong to
// the neighbour; they are sent home and added to its first/last owned rows.
int f2d_solver::reverse_exchange_add(float* buf) {
    if (!multi()) return F2D_OK;
    if (p2p) return reverse_exchange_p2p(buf);
    const size_t cnt = (size_t)H() * (size_t)g.pitch;
    F2D_NCCL(g_nccl.GroupStart());
    if (has_up()) {
        F2D_NCCL(g_nccl.Send(buf, cnt, kNcclFloat, rank - 1, comm, stream));
        F2D_NCCL(g_nccl.Recv(rx_up, cnt, kNcclFloat, rank - 1, comm, stream));
    }
    if (has_down()) {
        F2D_NCCL(g_nccl.Send(buf + (size_t)(g.rows - H()) * g.pitch, cnt, kNcclFloat, rank + 1, comm, stream));
        F2D_NCCL(g_nccl.Recv(rx_down, cnt, kNcclFloat, rank + 1, comm, stream));
    }
    F2D_NCCL(g_nccl.GroupEnd());
    if (has_up()) {
        launch_add_rows(g, buf, H(), H(), rx_up, stream);
        count();
    }
    if (has_down()) {
        launch_add_rows(g, buf, g.rows - 2 * H(), H(), rx_down, stream);
        count();
    }
    F2D_CUDA(cudaGetLastError());
    ++exchanges;
    xbytes += (uint64_t)cnt * sizeof(float);
    set_inv(buf, H());
    return F2D_OK;
}

// ------------------------------------------------------------- peer-to-peer transport (f2d_p2p.cu)
int f2d_solver::exchange_p2p(const float* const* bufs, int n) {
    const unsigned n4 = (unsigned)((size_t)H() * g.pitch / 4);
    const size_t halo_floats = (size_t)H() * g.pitch;
    XchgParams P;
    P.my_flags = flags;
    P.up_flags = has_up() ? peer_up.flags() : nullptr;
    P.down_flags = has_down() ? peer_down.flags() : nullptr;
    P.timeout_ns = p2p_timeout_ns;
    P.nseg = 0;
    for (int i = 0; i < n; ++i) {
        const float* b = bufs[i];
        const int idx = buffer_index(b);
        if (idx < 0 || idx >= nbuffers()) return fail(F2D_ERR_STATE, "exchange of a buffer outside the arena");
        if (has_up()) {  // my rows [H, 2H) -> the upper neighbour's bottom halo rows [rows_up - H, rows_up)
            float* dst = reinterpret_cast<float*>(peer_up.arena + (size_t)idx * peer_up.field_stride) +
                         (size_t)(peer_up.rows - H()) * g.pitch;
            P.seg[P.nseg++] = {reinterpret_cast<const float4*>(b + halo_floats), reinterpret_cast<float4*>(dst), n4};
        }
        if (has_down()) {  // my rows [rows - 2H, rows - H) -> the lower neighbour's top halo rows [0, H)
            float* dst = reinterpret_cast<float*>(peer_down.arena + (size_t)idx * peer_down.field_stride);
            P.seg[P.nseg++] = {reinterpret_cast<const float4*>(b + (size_t)(g.rows - 2 * H()) * g.pitch),
                               reinterpret_cast<float4*>(dst), n4};
        }
        if ((i + 1) % 4 == 0 || i == n - 1) {  // <= 4 buffers (8 segments) per kernel; chunking by BUFFER count so that
                                                // every rank (edge ranks have one neighbour) issues the same handshakes
            launch_halo_xchg(P, stream);
            count();
            ++exchanges;
            P.nseg = 0;
        }
    }
    F2D_CUDA(cudaGetLastError());
    for (int i = 0; i < n; ++i) set_inv(bufs[i], 0);
    xbytes += (uint64_t)n * halo_floats * sizeof(float);
    return F2D_OK;
}

int f2d_solver::reverse_exchange_p2p(float* buf) {
    const unsigned n4 = (unsigned)((size_t)H() * g.pitch / 4);
    XchgParams P;
    P.my_flags = flags;
    P.up_flags = has_up() ? peer_up.flags() : nullptr;
    P.down_flags = has_down() ? peer_down.flags() : nullptr;
    P.timeout_ns = p2p_timeout_ns;
    P.nseg = 0;
    // the partial sums in my halo rows go to the owner's landing zone: my top halo is the upper neighbour's
    // "from below" zone (rx_down there), my bottom halo the lower neighbour's "from above" zone (rx_up)
    if (has_up()) P.seg[P.nseg++] = {reinterpret_cast<const float4*>(buf), reinterpret_cast<float4*>(peer_up.rx_down()), n4};
    if (has_down())
        P.seg[P.nseg++] = {reinterpret_cast<const float4*>(buf + (size_t)(g.rows - H()) * g.pitch),
                           reinterpret_cast<float4*>(peer_down.rx_up()), n4};
    launch_halo_xchg(P, stream);
    count();
    if (has_up()) {
        launch_add_rows(g, buf, H(), H(), rx_up, stream);
        count();
    }
    if (has_down()) {
        launch_add_rows(g, buf, g.rows - 2 * H(), H(), rx_down, stream);
        count();
    }
    F2D_CUDA(cudaGetLastError());
    ++exchanges;
    xbytes += (uint64_t)n4 * 16u;
    set_inv(buf, H());
    return F2D_OK;
}

// =============================================================================== C ABI
extern "C" {

// ---- multi-GPU over peer memory: every rank exports one IPC handle for its arena; the host side gathers
// them (torch.distributed) and hands each rank its neighbours' handles.
F2D_API int f2d_p2p_export(f2d_solver* s, unsigned char* handle64, uint64_t* info4) {
    if (!s || !handle64 || !info4) return fail(F2D_ERR_INVALID, "NULL argument");
    F2D_CUDA(cudaSetDevice(s->device));
    cudaIpcMemHandle_t h;
    F2D_CUDA(cudaIpcGetMemHandle(&h, s->arena));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(handle64, &h, 64);
    info4[0] = s->field_stride;
    info4[1] = (uint64_t)s->g.rows;
    info4[2] = s->rx_bytes;
    info4[3] = (uint64_t)s->nbuffers();
    return F2D_OK;
}

F2D_API int f2d_p2p_connect(f2d_solver* s, int rank, int nranks, const unsigned char* up_handle64, const uint64_t* up_info4,
                            const unsigned char* down_handle64, const uint64_t* down_info4, int cfl_cells) {
    if (!s) return fail(F2D_ERR_INVALID, "NULL argument");
    if (nranks < 2 || rank < 0 || rank >= nranks) return fail(F2D_ERR_INVALID, "bad rank/nranks");
    if (s->cpu_sem()) return fail(F2D_ERR_INVALID, "F2D_SEM_CPU runs on one GPU only");
    if (s->cfg.halo == 0) return fail(F2D_ERR_INVALID, "a slab solver needs halo > 0");
    if (s->multi()) return fail(F2D_ERR_STATE, "communicator already initialised");
    if ((s->g.grow0 > 0) != (rank > 0) || (s->g.grow0 + s->g.rows < s->g.grows) != (rank < nranks - 1))
        return fail(F2D_ERR_INVALID, "slab position does not match rank (slabs are ordered by rank)");
    if (s->g.rows < 3 * (int)s->cfg.halo) return fail(F2D_ERR_INVALID, "slab thinner than 3 halos");
    if ((s->has_up() && (!up_handle64 || !up_info4)) || (s->has_down() && (!down_handle64 || !down_info4)))
        return fail(F2D_ERR_INVALID, "missing neighbour handle");
    F2D_CUDA(cudaSetDevice(s->device));
    auto open = [&](f2d_solver::PeerLink& L, const unsigned char* h64, const uint64_t* info) -> int {
        cudaIpcMemHandle_t h;
        memcpy(&h, h64, 64);
        void* p = nullptr;
        F2D_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        L.arena = static_cast<char*>(p);
        L.field_stride = (size_t)info[0];
        L.rows = (int)info[1];
        L.rx_bytes = (size_t)info[2];
        L.nbuf = (int)info[3];
        if (L.nbuf != s->nbuffers() || L.rx_bytes != s->rx_bytes) return fail(F2D_ERR_INVALID, "neighbour solver has a different configuration");
        return F2D_OK;
    };
    if (s->has_up()) F2D_TRY(open(s->peer_up, up_handle64, up_info4));
    if (s->has_down()) F2D_TRY(open(s->peer_down, down_handle64, down_info4));
    s->p2p = true;
    s->rank = rank;
    s->nranks = nranks;
    s->cfl_cells = cfl_cells > 0 ? std::min(cfl_cells, s->H() - 1) : s->H() - 1;
    s->drop_graphs();
    return F2D_OK;
}

// ---- multi-GPU bootstrap: rank 0 makes the id, the host side broadcasts it (torch.distributed),
// every rank calls f2d_comm_init (collective).
F2D_API int f2d_comm_unique_id(char* id128) {
    if (!id128) return fail(F2D_ERR_INVALID, "NULL argument");
    F2D_TRY(load_nccl());
    NcclId id;
    F2D_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(id128, id.internal, sizeof(id.internal));
    return F2D_OK;
}

F2D_API int f2d_comm_init(f2d_solver* s, const char* id128, int rank, int nranks, int cfl_cells) {
    if (!s || !id128) return fail(F2D_ERR_INVALID, "NULL argument");
    if (nranks < 2 || rank < 0 || rank >= nranks) return fail(F2D_ERR_INVALID, "bad rank/nranks");
    if (s->cpu_sem()) return fail(F2D_ERR_INVALID, "F2D_SEM_CPU runs on one GPU only");
    if (s->cfg.halo == 0) return fail(F2D_ERR_INVALID, "a slab solver needs halo > 0");
    if (s->comm) return fail(F2D_ERR_STATE, "communicator already initialised");
    if ((s->g.grow0 > 0) != (rank > 0) || (s->g.grow0 + s->g.rows < s->g.grows) != (rank < nranks - 1))
        return fail(F2D_ERR_INVALID, "slab position does not match rank (slabs are ordered by rank)");
    if (s->g.rows < 3 * (int)s->cfg.halo) return fail(F2D_ERR_INVALID, "slab thinner than 3 halos");
    F2D_CUDA(cudaSetDevice(s->device));
    F2D_TRY(load_nccl());
    NcclId id;
    memcpy(id.internal, id128, sizeof(id.internal));
    void* comm = nullptr;
    F2D_NCCL(g_nccl.CommInitRank(&comm, nranks, id, rank));
    s->comm = comm;
    s->rank = rank;
    s->nranks = nranks;
    s->cfl_cells = cfl_cells > 0 ? std::min(cfl_cells, s->H() - 1) : s->H() - 1;
    {
        // establish the P2P connections eagerly (outside any graph capture) with one exchange of each kind
        float* t = s->acquire();
        if (!t) return fail(F2D_ERR_STATE, "scratch pool exhausted");
        F2D_CUDA(cudaMemsetAsync(t, 0, s->field_bytes, s->stream));
        const float* one[1] = {t};
        int rc = s->exchange(one, 1);
        if (rc == F2D_OK) rc = s->reverse_exchange_add(t);
        s->release(t);
        if (rc != F2D_OK) return rc;
        F2D_CUDA(cudaStreamSynchronize(s->stream));
        s->exchanges = 0;
        s->inv_table.clear();
    }
    s->drop_graphs();  // a graph captured before had no exchanges in it
    return F2D_OK;
}

F2D_API int f2d_comm_stats(const f2d_solver* s, uint64_t* exchanges) {
    if (!s || !exchanges) return fail(F2D_ERR_INVALID, "NULL argument");
    *exchanges = s->exchanges;
    return F2D_OK;
}

F2D_API const char* f2d_last_error(void) { return g_err; }
F2D_API int f2d_abi_version(void) { return F2D_ABI_VERSION; }

F2D_API int f2d_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

F2D_API int f2d_config_default(f2d_config* cfg, uint32_t rows, uint32_t cols) {
    if (!cfg) return fail(F2D_ERR_INVALID, "cfg is NULL");
    memset(cfg, 0, sizeof(*cfg));
    cfg->struct_size = (uint32_t)sizeof(f2d_config);
    cfg->rows = rows;
    cfg->cols = cols;
    cfg->diffuse_iters = 15;
    cfg->project_iters = 20;
    cfg->smooth = 1;
    cfg->jacobi_mode = (cols % 4 == 0) ? F2D_JACOBI_STREAM : F2D_JACOBI_NAIVE;
    cfg->temporal_block = 0;
    cfg->divide_mode = F2D_DIV_F32_CORR;
    cfg->use_graph = 1;
    cfg->device = -1;
    cfg->global_rows = rows;
    cfg->row_offset = 0;
    cfg->halo = 0;
    cfg->stream = nullptr;
    return F2D_OK;
}

F2D_API int f2d_create(const f2d_config* cfg, f2d_solver** out) {
    if (!cfg || !out) return fail(F2D_ERR_INVALID, "NULL argument");
    *out = nullptr;
    if (cfg->struct_size != sizeof(f2d_config)) return fail(F2D_ERR_INVALID, "f2d_config.struct_size mismatch (ABI)");
    if (cfg->rows < 3 || cfg->cols < 3) return fail(F2D_ERR_INVALID, "grid must be at least 3x3");
    if (cfg->rows > (1u << 20) || cfg->cols > (1u << 20)) return fail(F2D_ERR_INVALID, "grid too large");
    const uint32_t grows = cfg->global_rows ? cfg->global_rows : cfg->rows;
    if (cfg->row_offset + cfg->rows > grows) return fail(F2D_ERR_INVALID, "slab exceeds global_rows");
    if (2 * (uint64_t)cfg->halo >= cfg->rows && cfg->halo != 0) return fail(F2D_ERR_INVALID, "halo too deep for slab");
    if ((uint64_t)grows * cfg->cols >= (1ull << 31) * 2) return fail(F2D_ERR_INVALID, "grid too large");
    if (cfg->jacobi_mode != F2D_JACOBI_NAIVE && cfg->jacobi_mode != F2D_JACOBI_STREAM)
        return fail(F2D_ERR_INVALID, "unknown jacobi_mode");
    if (cfg->divide_mode != F2D_DIV_F64 && cfg->divide_mode != F2D_DIV_F32_CORR)
        return fail(F2D_ERR_INVALID, "unknown divide_mode");
    if (cfg->semantics != F2D_SEM_GPU && cfg->semantics != F2D_SEM_CPU) return fail(F2D_ERR_INVALID, "unknown semantics");
    if (cfg->semantics == F2D_SEM_CPU && (cfg->halo != 0 || cfg->row_offset != 0 || grows != cfg->rows))
        return fail(F2D_ERR_INVALID, "F2D_SEM_CPU runs on one GPU only (the Gauss-Seidel wavefront is not split into slabs)");
    if (cfg->semantics == F2D_SEM_CPU && (uint64_t)cfg->rows * ((cfg->cols + 31u) / 32u * 32u) >= 0xfffffffeull)
        return fail(F2D_ERR_INVALID, "F2D_SEM_CPU: grid too large for 32-bit cell indices");

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(F2D_ERR_NO_DEVICE, "no CUDA device: libf2d has no CPU fallback");
    }
    int dev = cfg->device;
    if (dev < 0) F2D_CUDA(cudaGetDevice(&dev));
    if (dev >= ndev) return fail(F2D_ERR_INVALID, "device %d out of range (%d devices)", dev, ndev);
    F2D_CUDA(cudaSetDevice(dev));

    f2d_solver* s = new (std::nothrow) f2d_solver();
    if (!s) return fail(F2D_ERR_INVALID, "out of host memory");
    s->cfg = *cfg;
    s->cfg.global_rows = grows;
    s->device = dev;
    s->g.rows = (int)cfg->rows;
    s->g.cols = (int)cfg->cols;
    s->g.pitch = (int)((cfg->cols + 31u) / 32u * 32u);
    s->g.grow0 = (int)cfg->row_offset;
    s->g.grows = (int)grows;
    s->field_bytes = (size_t)s->g.pitch * (size_t)s->g.rows * sizeof(float);

    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) == cudaSuccess) s->sm_count = prop.multiProcessorCount;

    // sweeps fused per pass: 8 is fastest for both relaxations at 4096^2 and 16384^2
    // (profiles/tune_r01_run5_*.jsonl); the two can be set independently
    // grids up to ~1024^2 are launch/latency bound: shallower pipelines (less warm-up per chunk) win there
    // (profiles/tune_small_r01.log: 1024^2 K=40 0.304 ms at T=4 vs 0.342 at T=8; 256^2 K=20 0.107 vs 0.129)
    const int auto_T = ((uint64_t)grows * cfg->cols <= 1200ull * 1200ull) ? 4 : 8;
    if (s->cfg.temporal_block_diffuse == 0)
        s->cfg.temporal_block_diffuse =
            s->cfg.temporal_block ? s->cfg.temporal_block : (uint32_t)env_int("F2D_TEMPORAL_BLOCK_DIFFUSE", env_int("F2D_TEMPORAL_BLOCK", auto_T));
    if (s->cfg.temporal_block == 0) s->cfg.temporal_block = (uint32_t)env_int("F2D_TEMPORAL_BLOCK", auto_T);
    if (s->cfg.semantics == F2D_SEM_CPU) {
        s->cfg.jacobi_mode = F2D_JACOBI_NAIVE;  // unused: the relaxations are Gauss-Seidel wavefronts
        s->cfg.temporal_block = s->cfg.temporal_block_diffuse = 1;
    } else if (s->cfg.jacobi_mode == F2D_JACOBI_STREAM) {
        if (!stream_supported(s->g, (int)s->cfg.temporal_block) || !stream_supported(s->g, (int)s->cfg.temporal_block_diffuse)) {
            delete s;
            return fail(F2D_ERR_INVALID,
                        "F2D_JACOBI_STREAM needs cols %% 4 == 0 and temporal_block in {1,2,4,8}; use F2D_JACOBI_NAIVE");
        }
    } else {
        s->cfg.temporal_block = s->cfg.temporal_block_diffuse = 1;
    }
    s->tune.chunk_rows = env_int("F2D_STREAM_CHUNK_ROWS", 0);
    s->tune.warps_per_cta = env_int("F2D_STREAM_WARPS_PER_CTA", 0);
    s->tune.pdl = env_int("F2D_STREAM_PDL", 1);
    s->tune.min_blocks = env_int("F2D_STREAM_MIN_BLOCKS", 0);
    s->tune.min_chunk_mult = env_int("F2D_STREAM_MIN_CHUNK_MULT", 0);
    s->tune.edge_cost_pct = env_int("F2D_STREAM_EDGE_COST_PCT", 0);
    s->fuse_divergence = env_int("F2D_FUSE_DIVERGENCE", 1) != 0;
    s->fuse_sources = env_int("F2D_FUSE_SOURCES", 1) != 0;

    auto cleanup = [&](int rc) {
        f2d_destroy(s);
        return rc;
    };
    if (cfg->stream) {
        s->stream = (cudaStream_t)cfg->stream;
    } else {
        if (cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess)
            return cleanup(fail(F2D_ERR_CUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(cudaGetLastError())));
        s->own_stream = true;
    }
    // One arena for everything peers may touch: 6 state fields, the scratch pool, the two landing zones of
    // the reverse (scatter) exchange and a flag block.  One allocation == one CUDA IPC handle per rank, and a
    // buffer is identified across ranks by its index (every rank runs the same schedule on the same pool).
    const int ntemps = 13;  // deepest point: batched diffuse (3 x0 + 3x2 ping-pong + p/div views) + out_u/out_v of the solve() pipeline
    s->field_stride = (s->field_bytes + 511) / 512 * 512;
    s->rx_bytes = ((size_t)cfg->halo * s->g.pitch * sizeof(float) + 511) / 512 * 512;
    s->arena_bytes = (size_t)(6 + ntemps) * s->field_stride + 2 * s->rx_bytes + 4096;
    if (cudaMalloc(&s->arena, s->arena_bytes) != cudaSuccess)
        return cleanup(fail(F2D_ERR_CUDA, "cudaMalloc(%zu) failed: %s", s->arena_bytes, cudaGetErrorString(cudaGetLastError())));
    cudaMemsetAsync(s->arena, 0, s->arena_bytes, s->stream);
    for (int i = 0; i < 6; ++i) s->state[i] = reinterpret_cast<float*>(s->arena + (size_t)i * s->field_stride);
    for (int i = 0; i < ntemps; ++i) {
        s->temps.push_back(reinterpret_cast<float*>(s->arena + (size_t)(6 + i) * s->field_stride));
        s->temp_busy.push_back(0);
    }
    s->rx_up = reinterpret_cast<float*>(s->arena + (size_t)(6 + ntemps) * s->field_stride);
    s->rx_down = reinterpret_cast<float*>(s->arena + (size_t)(6 + ntemps) * s->field_stride + s->rx_bytes);
    s->flags = reinterpret_cast<unsigned*>(s->arena + (size_t)(6 + ntemps) * s->field_stride + 2 * s->rx_bytes);
    if (cudaMalloc(&s->oob_flag, sizeof(int)) != cudaSuccess)
        return cleanup(fail(F2D_ERR_CUDA, "cudaMalloc failed: %s", cudaGetErrorString(cudaGetLastError())));
    cudaMemsetAsync(s->oob_flag, 0, sizeof(int), s->stream);
    s->host_register = env_int("F2D_HOST_REGISTER", 0) != 0;
    if (s->cpu_sem()) {
        const uint32_t k = std::max(3u * s->cfg.diffuse_iters, s->cfg.project_iters);
        s->gs_flag_cap = gs_flag_words(s->g.rows, 1, (int)std::max(k, 1u));
        if (cudaMalloc(&s->gs_flags, s->gs_flag_cap * sizeof(unsigned)) != cudaSuccess ||
            cudaMalloc(&s->gs_aux, 4 * sizeof(unsigned)) != cudaSuccess)
            return cleanup(fail(F2D_ERR_CUDA, "cudaMalloc failed: %s", cudaGetErrorString(cudaGetLastError())));
        cudaMemsetAsync(s->gs_aux, 0, 4 * sizeof(unsigned), s->stream);
    }
    if (cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&s->ev_density, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&s->ev_copy, cudaEventDisableTiming) != cudaSuccess)
        return cleanup(fail(F2D_ERR_CUDA, "stream/event creation failed: %s", cudaGetErrorString(cudaGetLastError())));
    if (cudaStreamCreateWithFlags(&s->up_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&s->ev_in[0], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&s->ev_in[1], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&s->ev_in[2], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&s->ev_vel, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&s->ev_fence, cudaEventDisableTiming) != cudaSuccess)
        return cleanup(fail(F2D_ERR_CUDA, "stream/event creation failed: %s", cudaGetErrorString(cudaGetLastError())));
    s->p2p_timeout_ns = (unsigned long long)std::max(1, env_int("F2D_P2P_TIMEOUT_MS", 30000)) * 1000000ull;
    s->host_pipeline = env_int("F2D_HOST_PIPELINE", 1) != 0;
    s->host_pipeline_min_bytes = (size_t)env_int("F2D_HOST_PIPELINE_MIN_BYTES", 1 << 20);
    if (cudaEventCreate(&s->ev0) != cudaSuccess || cudaEventCreate(&s->ev1) != cudaSuccess)
        return cleanup(fail(F2D_ERR_CUDA, "cudaEventCreate failed: %s", cudaGetErrorString(cudaGetLastError())));
    if (cudaStreamSynchronize(s->stream) != cudaSuccess)
        return cleanup(fail(F2D_ERR_CUDA, "initialisation failed: %s", cudaGetErrorString(cudaGetLastError())));
    *out = s;
    return F2D_OK;
}

F2D_API void f2d_destroy(f2d_solver* s) {
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    if (s->graph_exec) cudaGraphExecDestroy(s->graph_exec);
    for (auto& e : s->host_exec)
        if (e) cudaGraphExecDestroy(e);
    if (s->up_stream) {
        cudaStreamSynchronize(s->up_stream);
        cudaStreamDestroy(s->up_stream);
    }
    for (auto& e : s->ev_in)
        if (e) cudaEventDestroy(e);
    if (s->ev_vel) cudaEventDestroy(s->ev_vel);
    if (s->ev_fence) cudaEventDestroy(s->ev_fence);
    for (int i = 0; i < 6; ++i)
        s->state[i] = nullptr;
    if (s->arena) cudaFree(s->arena);
    if (s->oob_flag) cudaFree(s->oob_flag);
    if (s->render_buf) cudaFree(s->render_buf);
    if (s->gs_flags) cudaFree(s->gs_flags);
    if (s->gs_aux) cudaFree(s->gs_aux);
    if (s->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(s->comm);
    if (s->peer_up.arena) cudaIpcCloseMemHandle(s->peer_up.arena);
    if (s->peer_down.arena) cudaIpcCloseMemHandle(s->peer_down.arena);
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    if (s->copy_stream) {
        cudaStreamSynchronize(s->copy_stream);
        cudaStreamDestroy(s->copy_stream);
    }
    if (s->ev_density) cudaEventDestroy(s->ev_density);
    if (s->ev_copy) cudaEventDestroy(s->ev_copy);
    for (auto& r : s->registered)
        if (cudaHostUnregister(r.first) != cudaSuccess) cudaGetLastError();
    if (s->own_stream && s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

#define F2D_NEED(s)                                              \
    do {                                                         \
        if (!(s)) return fail(F2D_ERR_INVALID, "solver is NULL"); \
        F2D_CUDA(cudaSetDevice((s)->device));                    \
    } while (0)

F2D_API int f2d_pin_host(f2d_solver* s, const void* host, size_t bytes) {
    F2D_NEED(s);
    if (!host || bytes == 0) return fail(F2D_ERR_INVALID, "NULL host pointer or empty range");
    for (auto& r : s->registered)
        if (r.first == host) return r.second >= bytes ? F2D_OK : fail(F2D_ERR_STATE, "range already pinned with a smaller size");
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, host) == cudaSuccess && at.type != cudaMemoryTypeUnregistered) return F2D_OK;  // already page-locked
    cudaGetLastError();
    F2D_CUDA(cudaHostRegister(const_cast<void*>(host), bytes, cudaHostRegisterDefault));
    s->registered.emplace_back(const_cast<void*>(host), bytes);
    return F2D_OK;
}

F2D_API int f2d_unpin_host(f2d_solver* s, const void* host) {
    F2D_NEED(s);
    for (size_t i = 0; i < s->registered.size(); ++i)
        if (s->registered[i].first == host) {
            // no copy of this solver may still be using the range
            F2D_CUDA(cudaStreamSynchronize(s->stream));
            F2D_CUDA(cudaStreamSynchronize(s->up_stream));
            F2D_CUDA(cudaStreamSynchronize(s->copy_stream));
            F2D_CUDA(cudaHostUnregister(const_cast<void*>(host)));
            s->registered.erase(s->registered.begin() + (ptrdiff_t)i);
            return F2D_OK;
        }
    return F2D_OK;  // not pinned by this solver (e.g. it was page-locked already): nothing to undo
}

F2D_API int f2d_comm_bytes(const f2d_solver* s, uint64_t* to_up, uint64_t* to_down) {
    if (!s || !to_up || !to_down) return fail(F2D_ERR_INVALID, "NULL argument");
    *to_up = s->has_up() ? s->xbytes : 0;
    *to_down = s->has_down() ? s->xbytes : 0;
    return F2D_OK;
}

F2D_API int f2d_upload_field(f2d_solver* s, int field, const float* host) {
    F2D_NEED(s);
    if (field < 0 || field >= 6 || !host) return fail(F2D_ERR_INVALID, "bad field or NULL host pointer");
    F2D_TRY(s->h2d(s->state[field], host));
    F2D_CUDA(cudaStreamSynchronize(s->stream));
    return F2D_OK;
}

F2D_API int f2d_download_field(f2d_solver* s, int field, float* host) {
    F2D_NEED(s);
    float* p = s->field_ptr(field);
    if (!p || !host) return fail(F2D_ERR_INVALID, "field %d not available or NULL host pointer", field);
    F2D_TRY(s->d2h(host, p));
    F2D_CUDA(cudaStreamSynchronize(s->stream));
    return F2D_OK;
}

F2D_API int f2d_upload(f2d_solver* s, const float* density, const float* u, const float* v) {
    F2D_NEED(s);
    if (density) F2D_TRY(s->h2d(s->state[F2D_FIELD_DENSITY], density));
    if (u) F2D_TRY(s->h2d(s->state[F2D_FIELD_U], u));
    if (v) F2D_TRY(s->h2d(s->state[F2D_FIELD_V], v));
    F2D_CUDA(cudaStreamSynchronize(s->stream));
    return F2D_OK;
}

F2D_API int f2d_set_sources(f2d_solver* s, const float* sd, const float* su, const float* sv) {
    F2D_NEED(s);
    if (sd) F2D_TRY(s->h2d(s->state[F2D_FIELD_DENSITY_SOURCE], sd));
    if (su) F2D_TRY(s->h2d(s->state[F2D_FIELD_U_SOURCE], su));
    if (sv) F2D_TRY(s->h2d(s->state[F2D_FIELD_V_SOURCE], sv));
    F2D_CUDA(cudaStreamSynchronize(s->stream));
    return F2D_OK;
}

F2D_API int f2d_clear_sources(f2d_solver* s) {
    F2D_NEED(s);
    for (int f = F2D_FIELD_DENSITY_SOURCE; f <= F2D_FIELD_V_SOURCE; ++f)
        F2D_CUDA(cudaMemsetAsync(s->state[f], 0, s->field_bytes, s->stream));
    return F2D_OK;
}

F2D_API int f2d_download(f2d_solver* s, float* density, float* u, float* v) {
    F2D_NEED(s);
    if (density) F2D_TRY(s->d2h(density, s->state[F2D_FIELD_DENSITY]));
    if (u) F2D_TRY(s->d2h(u, s->state[F2D_FIELD_U]));
    if (v) F2D_TRY(s->d2h(v, s->state[F2D_FIELD_V]));
    F2D_CUDA(cudaStreamSynchronize(s->stream));
    return F2D_OK;
}

F2D_API int f2d_step(f2d_solver* s, float diffusion_rate, float viscosity, float dt, uint32_t nsteps) {
    F2D_NEED(s);
    return s->step(diffusion_rate, viscosity, dt, nsteps);
}

F2D_API int f2d_step_timed(f2d_solver* s, float diffusion_rate, float viscosity, float dt, uint32_t nsteps, float* elapsed_ms) {
    F2D_NEED(s);
    if (!elapsed_ms) return fail(F2D_ERR_INVALID, "elapsed_ms is NULL");
    if (s->cfg.use_graph) F2D_TRY(s->ensure_graph(diffusion_rate, viscosity, dt));  // keep capture out of the timing
    F2D_CUDA(cudaStreamSynchronize(s->stream));
    F2D_CUDA(cudaEventRecord(s->ev0, s->stream));
    F2D_TRY(s->step(diffusion_rate, viscosity, dt, nsteps));
    F2D_CUDA(cudaEventRecord(s->ev1, s->stream));
    F2D_CUDA(cudaEventSynchronize(s->ev1));
    F2D_CUDA(cudaEventElapsedTime(elapsed_ms, s->ev0, s->ev1));
    return F2D_OK;
}

F2D_API int f2d_sync(f2d_solver* s) {
    F2D_NEED(s);
    F2D_CUDA(cudaStreamSynchronize(s->stream));
    int oob = 0;
    F2D_CUDA(cudaMemcpy(&oob, s->oob_flag, sizeof(int), cudaMemcpyDeviceToHost));
    if (s->p2p) {
        unsigned err = 0;
        F2D_CUDA(cudaMemcpy(&err, s->flags + 2, sizeof(unsigned), cudaMemcpyDeviceToHost));
        if (err) {
            cudaMemset(s->flags + 2, 0, sizeof(unsigned));  // report once; later exchanges wait normally again
            return fail(F2D_ERR_STATE, "halo exchange timed out waiting for a neighbour GPU (state invalid since then)");
        }
    }
    if (s->gs_aux) {
        unsigned err = 0;
        F2D_CUDA(cudaMemcpy(&err, s->gs_aux, sizeof(unsigned), cudaMemcpyDeviceToHost));
        if (err) {
            cudaMemset(s->gs_aux, 0, sizeof(unsigned));
            return fail(F2D_ERR_STATE, "Gauss-Seidel wavefront: a tile waited for its neighbours for too long (results invalid)");
        }
    }
    if (oob) {
        cudaMemset(s->oob_flag, 0, sizeof(int));
        return fail(F2D_ERR_STATE, "advection left the rows this slab holds valid: the displacement of a step exceeded cfl_cells "
                                   "(at most halo - 1 rows); the state is invalid from that step on");
    }
    return F2D_OK;
}

F2D_API int f2d_solve_host(f2d_solver* s, float* density, const float* density_source, float diffusion_rate, float* u, float* v,
                   const float* u_source, const float* v_source, float viscosity, float dt) {
    F2D_NEED(s);
    if (!density || !density_source || !u || !v || !u_source || !v_source) return fail(F2D_ERR_INVALID, "NULL grid pointer");
    const size_t host_bytes = (size_t)s->g.rows * s->g.cols * sizeof(float);
    const void* hosts[6] = {density, u, v, density_source, u_source, v_source};
    for (const void* h : hosts) s->pin_host(h, host_bytes);
    if (s->host_pipeline && !s->cpu_sem() && s->pipeline_decision_bytes() >= s->host_pipeline_min_bytes) {
        F2D_TRY(s->solve_host_pipelined(density, density_source, diffusion_rate, u, v, u_source, v_source, viscosity, dt));
        return f2d_sync(s);
    }
    // upload (gpu.cu:232-234 and the source uploads of :281) -> one step -> download (:255-257)
    F2D_TRY(s->h2d(s->state[F2D_FIELD_DENSITY], density));
    F2D_TRY(s->h2d(s->state[F2D_FIELD_DENSITY_SOURCE], density_source));
    F2D_TRY(s->h2d(s->state[F2D_FIELD_U], u));
    F2D_TRY(s->h2d(s->state[F2D_FIELD_V], v));
    F2D_TRY(s->h2d(s->state[F2D_FIELD_U_SOURCE], u_source));
    F2D_TRY(s->h2d(s->state[F2D_FIELD_V_SOURCE], v_source));
    F2D_TRY(s->step(diffusion_rate, viscosity, dt, 1));
    // density is final once ev_density fires (after the scatter + smooth): download it on the copy stream
    // while the two projections and the advection are still running
    F2D_CUDA(cudaStreamWaitEvent(s->copy_stream, s->ev_density, 0));
    F2D_TRY(s->d2h_on(density, s->state[F2D_FIELD_DENSITY], s->copy_stream));
    F2D_CUDA(cudaEventRecord(s->ev_copy, s->copy_stream));
    F2D_TRY(s->d2h(u, s->state[F2D_FIELD_U]));
    F2D_TRY(s->d2h(v, s->state[F2D_FIELD_V]));
    F2D_CUDA(cudaStreamWaitEvent(s->stream, s->ev_copy, 0));  // the next upload must not overtake the download
    F2D_CUDA(cudaStreamSynchronize(s->copy_stream));
    return f2d_sync(s);
}

// ------------------------------------------------------------------------------ stages
static int copy_back(f2d_solver* s, float* dst, const float* src) {
    if (dst == src) return F2D_OK;
    F2D_CUDA(cudaMemcpyAsync(dst, src, s->field_bytes, cudaMemcpyDeviceToDevice, s->stream));
    return F2D_OK;
}

F2D_API int f2d_stage_set_bnd(f2d_solver* s, int field, int kind) {
    F2D_NEED(s);
    if (field < 0 || field >= 6) return fail(F2D_ERR_INVALID, "bad field");
    launch_set_bnd_inplace(s->g, s->state[field], kind, s->stream);
    s->count();
    F2D_CUDA(cudaGetLastError());
    if (s->cpu_sem()) return s->corners({s->state[field]});  // cpp:44-47
    return F2D_OK;
}

F2D_API int f2d_stage_add_sources(f2d_solver* s, int field, float dt) {
    F2D_NEED(s);
    if (field < 0 || field > F2D_FIELD_V) return fail(F2D_ERR_INVALID, "bad field");
    AddSourceBatch ab;
    ab.n = 1;
    ab.f[0] = s->state[field];
    ab.o[0] = s->state[field];
    ab.s[0] = s->state[field + 3];
    if (s->cpu_sem())
        launch_add_sources_nofma(s->g, ab, dt, s->stream);  // cpp:85-93
    else
        launch_add_sources(s->g, ab, dt, s->stream);
    s->count();
    F2D_CUDA(cudaGetLastError());
    return F2D_OK;
}

F2D_API int f2d_stage_diffuse(f2d_solver* s, int field, int kind, float rate, float dt, uint32_t iters) {
    F2D_NEED(s);
    if (field < 0 || field > F2D_FIELD_V) return fail(F2D_ERR_INVALID, "bad field");
    if (s->cpu_sem()) {  // cpp:95-114
        float* x[1] = {s->state[field]};
        const int kinds[1] = {kind};
        const float rates[1] = {rate};
        return s->diffuse_cpu(1, x, kinds, rates, dt, iters);
    }
    const float* in[1] = {s->state[field]};
    const float* rhs[1] = {s->state[field]};
    const int kinds[1] = {kind};
    const DiffuseCoef kc[1] = {s->diffuse_coef(rate, dt)};
    const float* out[1];
    F2D_TRY(s->relax(1, in, rhs, kinds, kc, true, iters, out));
    F2D_TRY(copy_back(s, s->state[field], out[0]));
    if (out[0] != s->state[field]) s->release(out[0]);
    return F2D_OK;
}

F2D_API int f2d_stage_smooth(f2d_solver* s) {
    F2D_NEED(s);
    // smooth_kernel alone (gpu.cu:314-323): the state already carries its boundary values, and the
    // fused kernel's edge rule reproduces them only if set_bnd was applied, so run it on a copy whose
    // edges are then restored from the input.
    float* t = s->acquire();
    float* d = s->state[F2D_FIELD_DENSITY];
    if (!t) return fail(F2D_ERR_STATE, "scratch pool exhausted");
    F2D_TRY(copy_back(s, t, d));
    launch_smooth_plain(s->g, t, d, s->stream);
    s->count();
    s->release(t);
    F2D_CUDA(cudaGetLastError());
    return F2D_OK;
}

F2D_API int f2d_stage_advect_density(f2d_solver* s, float dt) {
    F2D_NEED(s);
    float* d = s->state[F2D_FIELD_DENSITY];
    if (s->cpu_sem()) return s->advect_density_cpu(d, s->state[F2D_FIELD_U], s->state[F2D_FIELD_V], dt, false);  // cpp:116-177
    float* sc = s->acquire();
    if (!sc) return fail(F2D_ERR_STATE, "scratch pool exhausted");
    F2D_CUDA(cudaMemsetAsync(sc, 0, s->field_bytes, s->stream));
    launch_scatter_density(s->g, d, s->state[F2D_FIELD_U], s->state[F2D_FIELD_V], sc, s->dt0(dt), s->own_begin(),
                           s->own_end(), s->oob_flag, s->stream);
    s->count();
    launch_smooth_bnd(s->g, sc, d, false, s->stream);  // out-of-place set_boundary_continuous (gpu.cu:355)
    s->count();
    s->release(sc);
    F2D_CUDA(cudaGetLastError());
    return F2D_OK;
}

F2D_API int f2d_stage_advect_velocity(f2d_solver* s, float dt) {
    F2D_NEED(s);
    float *u = s->state[F2D_FIELD_U], *v = s->state[F2D_FIELD_V];
    float *u0 = s->acquire(), *v0 = s->acquire();
    if (!u0 || !v0) return fail(F2D_ERR_STATE, "scratch pool exhausted");
    F2D_TRY(copy_back(s, u0, u));  // gpu.cu:248-249
    F2D_TRY(copy_back(s, v0, v));
    if (s->cpu_sem())
        launch_advect_velocity_nofma(s->g, u0, v0, u, v, s->dt0_cpu(dt), s->stream);  // cpp:26-29
    else
        launch_advect_velocity(s->g, u0, v0, u, v, s->dt0(dt), s->own_begin(), s->own_end(), 0, s->g.rows, s->oob_flag, s->stream);
    s->count();
    s->release(u0);
    s->release(v0);
    F2D_CUDA(cudaGetLastError());
    if (s->cpu_sem()) return s->corners({u, v});
    return F2D_OK;
}

F2D_API int f2d_stage_project(f2d_solver* s, uint32_t iters) {
    F2D_NEED(s);
    float *u = s->state[F2D_FIELD_U], *v = s->state[F2D_FIELD_V];
    float *u0 = s->acquire(), *v0 = s->acquire();
    if (!u0 || !v0) return fail(F2D_ERR_STATE, "scratch pool exhausted");
    F2D_TRY(copy_back(s, u0, u));
    F2D_TRY(copy_back(s, v0, v));
    int rc = s->cpu_sem() ? s->project_cpu(u0, v0, u, v, iters) : s->project(u0, v0, u, v, iters);
    s->release(u0);
    s->release(v0);
    return rc;
}

F2D_API int f2d_bench_jacobi(f2d_solver* s, int diffuse_like, uint32_t iters, uint32_t reps, float* elapsed_ms) {
    F2D_NEED(s);
    if (!elapsed_ms || reps == 0) return fail(F2D_ERR_INVALID, "bad arguments");
    if (s->cpu_sem()) {
        // the Gauss-Seidel wavefront on a scratch copy: x = copy of u (diffuse) or zero (pressure), rhs = density
        float* x = s->acquire();
        if (!x) return fail(F2D_ERR_STATE, "scratch pool exhausted");
        if (diffuse_like)
            F2D_CUDA(cudaMemcpyAsync(x, s->state[F2D_FIELD_U], s->field_bytes, cudaMemcpyDeviceToDevice, s->stream));
        else
            F2D_CUDA(cudaMemsetAsync(x, 0, s->field_bytes, s->stream));
        float* xs[1] = {x};
        const float* rhs1[1] = {s->state[F2D_FIELD_DENSITY]};
        const int kinds1[1] = {F2D_BND_CONTINUOUS};
        const float a1[1] = {s->diffuse_coef(1e-6f, 0.02f).a};
        int rc = s->relax_gs(1, xs, rhs1, kinds1, diffuse_like ? a1 : nullptr, diffuse_like != 0, iters);  // warm-up
        if (rc == F2D_OK && cudaStreamSynchronize(s->stream) == cudaSuccess) {
            cudaEventRecord(s->ev0, s->stream);
            for (uint32_t r = 0; r < reps && rc == F2D_OK; ++r)
                rc = s->relax_gs(1, xs, rhs1, kinds1, diffuse_like ? a1 : nullptr, diffuse_like != 0, iters);
            cudaEventRecord(s->ev1, s->stream);
            if (rc == F2D_OK && (cudaEventSynchronize(s->ev1) != cudaSuccess || cudaEventElapsedTime(elapsed_ms, s->ev0, s->ev1) != cudaSuccess))
                rc = fail(F2D_ERR_CUDA, "timing the Gauss-Seidel wavefront failed: %s", cudaGetErrorString(cudaGetLastError()));
        }
        s->release(x);
        return rc;
    }
    // scratch problem: rhs = density state, start = u state (diffuse) or zero (pressure)
    const float* in[1] = {diffuse_like ? s->state[F2D_FIELD_U] : nullptr};
    const float* rhs[1] = {s->state[F2D_FIELD_DENSITY]};
    const int kinds[1] = {F2D_BND_CONTINUOUS};
    const DiffuseCoef kc[1] = {s->diffuse_coef(1e-6f, 0.02f)};
    const float* out[1];
    F2D_TRY(s->relax(1, in, rhs, kinds, kc, diffuse_like != 0, iters, out));  // warm-up
    if (out[0] != in[0]) s->release(out[0]);
    F2D_CUDA(cudaStreamSynchronize(s->stream));
    F2D_CUDA(cudaEventRecord(s->ev0, s->stream));
    for (uint32_t r = 0; r < reps; ++r) {
        F2D_TRY(s->relax(1, in, rhs, kinds, kc, diffuse_like != 0, iters, out));
        if (out[0] != in[0]) s->release(out[0]);
    }
    F2D_CUDA(cudaEventRecord(s->ev1, s->stream));
    F2D_CUDA(cudaEventSynchronize(s->ev1));
    F2D_CUDA(cudaEventElapsedTime(elapsed_ms, s->ev0, s->ev1));
    return F2D_OK;
}

// ------------------------------------------------------------------------------ renderers
static int render_common(f2d_solver* s, size_t bytes, void** scratch) {
    if (s->render_bytes < bytes) {
        if (s->render_buf) cudaFree(s->render_buf);
        s->render_buf = nullptr;
        s->render_bytes = 0;
        F2D_CUDA(cudaMalloc(&s->render_buf, bytes));
        s->render_bytes = bytes;
    }
    *scratch = s->render_buf;
    return F2D_OK;
}

F2D_API int f2d_render_density_rgba(f2d_solver* s, float mult_r, float mult_g, float mult_b, unsigned char* host_rgba) {
    F2D_NEED(s);
    if (!host_rgba) return fail(F2D_ERR_INVALID, "NULL argument");
    const size_t bytes = (size_t)s->g.rows * s->g.cols * 4;
    void* buf = nullptr;
    F2D_TRY(render_common(s, bytes, &buf));
    launch_density_to_rgba(s->g, s->state[F2D_FIELD_DENSITY], buf, mult_r, mult_g, mult_b, s->stream);
    s->count();
    F2D_CUDA(cudaGetLastError());
    F2D_CUDA(cudaMemcpyAsync(host_rgba, buf, bytes, cudaMemcpyDeviceToHost, s->stream));
    F2D_CUDA(cudaStreamSynchronize(s->stream));
    return F2D_OK;
}

F2D_API int f2d_render_velocity_lines(f2d_solver* s, float horizontal_scale, float vertical_scale, float* host_lines) {
    F2D_NEED(s);
    if (!host_lines) return fail(F2D_ERR_INVALID, "NULL argument");
    const size_t bytes = (size_t)s->g.rows * s->g.cols * 4 * sizeof(float);
    void* buf = nullptr;
    F2D_TRY(render_common(s, bytes, &buf));
    launch_velocity_to_lines(s->g, s->state[F2D_FIELD_U], s->state[F2D_FIELD_V], buf, horizontal_scale, vertical_scale,
                             sqrtf((float)s->global_cells()), s->stream);
    s->count();
    F2D_CUDA(cudaGetLastError());
    F2D_CUDA(cudaMemcpyAsync(host_lines, buf, bytes, cudaMemcpyDeviceToHost, s->stream));
    F2D_CUDA(cudaStreamSynchronize(s->stream));
    return F2D_OK;
}

F2D_API int f2d_launch_count(const f2d_solver* s, uint64_t* launches) {
    if (!s || !launches) return fail(F2D_ERR_INVALID, "NULL argument");
    *launches = s->launches;
    return F2D_OK;
}

F2D_API int f2d_field_ptr(f2d_solver* s, int field, void** device_ptr, size_t* pitch_elems) {
    if (!s || !device_ptr) return fail(F2D_ERR_INVALID, "NULL argument");
    float* p = s->field_ptr(field);
    if (!p) return fail(F2D_ERR_INVALID, "field %d not available", field);
    *device_ptr = p;
    if (pitch_elems) *pitch_elems = (size_t)s->g.pitch;
    return F2D_OK;
}

F2D_API int f2d_get_config(const f2d_solver* s, f2d_config* out) {
    if (!s || !out) return fail(F2D_ERR_INVALID, "NULL argument");
    *out = s->cfg;
    return F2D_OK;
}

F2D_API int f2d_get_stream(const f2d_solver* s, void** stream) {
    if (!s || !stream) return fail(F2D_ERR_INVALID, "NULL argument");
    *stream = (void*)s->stream;
    return F2D_OK;
}

}  // extern "C"
