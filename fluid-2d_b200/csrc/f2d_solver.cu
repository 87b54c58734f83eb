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
#include <new>
#include <vector>

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
    StreamTuning tune = {0, 0, 0};

    cudaEvent_t ev0 = nullptr, ev1 = nullptr;

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
    void release_all() {
        for (size_t i = 0; i < temps.size(); ++i) temp_busy[i] = 0;
    }
    bool is_temp(const float* p) const {
        for (size_t i = 0; i < temps.size(); ++i)
            if (temps[i] == p) return true;
        return false;
    }

    // ---- scalars of the reference (computed on the host exactly as it does) --------------
    size_t global_cells() const { return (size_t)g.grows * (size_t)g.cols; }
    DiffuseCoef diffuse_coef(float rate, float dt) const {
        DiffuseCoef k;
        float t = dt * (float)global_cells();  // src/fluid_solver_gpu.cu:79, left to right in fp32
        k.a = t * rate;
        k.c = 1.0 + 4.0 * (double)k.a;  // gpu.cu:82
        k.rc = (float)(1.0 / k.c);
        k.ch = (float)k.c;
        k.cl = (float)(k.c - (double)k.ch);
        return k;
    }
    float dt0(float dt) const { return (float)(std::sqrt((double)global_cells()) * (double)dt); }  // gpu.cu:334
    float h() const { return 1.0f / sqrtf((float)global_cells()); }                                  // gpu.cu:367

    // ---- building blocks -----------------------------------------------------------------
    void count(int n = 1) { launches += (uint64_t)n; }

    // K relaxation sweeps for n problems.  in[i] == nullptr means a zero start (pressure).
    // out[i] receives the buffer holding the final iterate: a pool buffer the caller must
    // release, or in[i] itself when K == 0.
    int relax(int n, const float* const* in, const float* const* rhs, const int* kinds, const DiffuseCoef* coefs,
              bool diffuse, uint32_t K, const float** out) {
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
            pong[i] = (K > 1 || true) ? acquire() : nullptr;
            if (!ping[i] || !pong[i]) return fail(F2D_ERR_STATE, "scratch pool exhausted");
        }
        const bool stream_mode = (cfg.jacobi_mode == F2D_JACOBI_STREAM);
        uint32_t left = K;
        int flip = 0;
        while (left > 0) {
            uint32_t T = 1;
            if (stream_mode) {
                T = cfg.temporal_block;
                while (T > left) T >>= 1;
            }
            RelaxBatch b;
            b.n = n;
            for (int i = 0; i < n; ++i) {
                b.f[i].prev = cur[i];
                b.f[i].rhs = rhs[i];
                b.f[i].next = flip ? pong[i] : ping[i];
                b.f[i].kind = kinds[i];
                b.f[i].coef = coefs ? coefs[i] : DiffuseCoef{0.f, 0.f, 0.f, 0.f, 1.0};
            }
            if (stream_mode)
                launch_jacobi_stream(g, b, diffuse, (int)cfg.divide_mode, (int)T, (int)T, tune, sm_count, stream);
            else
                launch_jacobi_naive(g, b, diffuse, (int)cfg.divide_mode, stream);
            count();
            for (int i = 0; i < n; ++i) cur[i] = b.f[i].next;
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
        launch_divergence(g, u_in, v_in, dv, h(), stream);
        count();
        const float* pin[1] = {nullptr};
        const float* prhs[1] = {dv};
        const int kind[1] = {F2D_BND_CONTINUOUS};
        const float* pout[1];
        F2D_TRY(relax(1, pin, prhs, kind, nullptr, false, K, pout));
        launch_gradient(g, pout[0], u_in, v_in, u_out, v_out, h(), stream);
        count();
        F2D_CUDA(cudaGetLastError());
        // keep p / div alive as debug views until the next project
        last_div = dv;
        last_p = const_cast<float*>(pout[0]);
        return F2D_OK;
    }

    // One full solve() step on the device-resident state (order of gpu.cu:236-252).
    int enqueue_step(float diffusion_rate, float viscosity, float dt) {
        float *d = state[F2D_FIELD_DENSITY], *u = state[F2D_FIELD_U], *v = state[F2D_FIELD_V];
        // ---------------- density chain (uses the PRE-step u, v) ----------------
        {
            AddSourceBatch ab;
            ab.n = 1;
            ab.f[0] = d;
            ab.s[0] = state[F2D_FIELD_DENSITY_SOURCE];
            launch_add_sources(g, ab, dt, stream);  // gpu.cu:237
            count();
            const float* in[1] = {d};
            const float* rhs[1] = {d};  // x0 == the field after add_sources (gpu.cu:300)
            const int kind[1] = {F2D_BND_CONTINUOUS};
            const DiffuseCoef kc[1] = {diffuse_coef(diffusion_rate, dt)};
            const float* dd[1];
            F2D_TRY(relax(1, in, rhs, kind, kc, true, cfg.diffuse_iters, dd));  // gpu.cu:238
            float* sc = acquire();
            if (!sc) return fail(F2D_ERR_STATE, "scratch pool exhausted");
            F2D_CUDA(cudaMemsetAsync(sc, 0, field_bytes, stream));  // gpu.cu:337
            launch_scatter_density(g, dd[0], u, v, sc, dt0(dt), own_begin(), own_end(), oob_flag, stream);  // gpu.cu:239
            count();
            if (dd[0] != d) release(dd[0]);
            launch_smooth_bnd(g, sc, d, cfg.smooth != 0, stream);  // gpu.cu:355 + :240
            count();
            release(sc);
        }
        // ---------------- velocity chain ----------------
        {
            AddSourceBatch ab;
            ab.n = 2;
            ab.f[0] = u;
            ab.s[0] = state[F2D_FIELD_U_SOURCE];
            ab.f[1] = v;
            ab.s[1] = state[F2D_FIELD_V_SOURCE];
            launch_add_sources(g, ab, dt, stream);  // gpu.cu:243-244
            count();
            const float* in[2] = {u, v};
            const float* rhs[2] = {u, v};
            const int kind[2] = {F2D_BND_OPPOSITE_HORIZONTAL, F2D_BND_OPPOSITE_VERTICAL};
            const DiffuseCoef kc[2] = {diffuse_coef(viscosity, dt), diffuse_coef(viscosity, dt)};
            const float* uv1[2];
            F2D_TRY(relax(2, in, rhs, kind, kc, true, cfg.diffuse_iters, uv1));  // gpu.cu:245-246
            float *u2 = acquire(), *v2 = acquire();
            if (!u2 || !v2) return fail(F2D_ERR_STATE, "scratch pool exhausted");
            F2D_TRY(project(uv1[0], uv1[1], u2, v2, cfg.project_iters));  // gpu.cu:247
            // advect both components by (U0,V0) = (u2,v2) (gpu.cu:248-251)
            float *u3, *v3;
            if (uv1[0] != u) {
                u3 = const_cast<float*>(uv1[0]);
                v3 = const_cast<float*>(uv1[1]);
            } else {
                u3 = acquire();
                v3 = acquire();
                if (!u3 || !v3) return fail(F2D_ERR_STATE, "scratch pool exhausted");
            }
            launch_advect_velocity(g, u2, v2, u3, v3, dt0(dt), stream);
            count();
            release(u2);
            release(v2);
            F2D_TRY(project(u3, v3, u, v, cfg.project_iters));  // gpu.cu:252, result lands in the state buffers
            release(u3);
            release(v3);
        }
        F2D_CUDA(cudaGetLastError());
        return F2D_OK;
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
        const uint64_t before = launches;
        cudaGraph_t graph = nullptr;
        F2D_CUDA(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
        int rc = enqueue_step(diffusion_rate, viscosity, dt);
        cudaError_t ce = cudaStreamEndCapture(stream, &graph);
        if (rc != F2D_OK) {
            if (graph) cudaGraphDestroy(graph);
            return rc;
        }
        if (ce != cudaSuccess) return fail(F2D_ERR_CUDA, "cudaStreamEndCapture failed: %s", cudaGetErrorString(ce));
        graph_kernels = launches - before;
        launches = before;  // capture launched nothing yet
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
            }
        } else {
            for (uint32_t s = 0; s < nsteps; ++s) F2D_TRY(enqueue_step(diffusion_rate, viscosity, dt));
        }
        return F2D_OK;
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

// =============================================================================== C ABI
extern "C" {

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

    if (s->cfg.temporal_block == 0) s->cfg.temporal_block = (uint32_t)env_int("F2D_TEMPORAL_BLOCK", 8);
    if (s->cfg.jacobi_mode == F2D_JACOBI_STREAM) {
        if (!stream_supported(s->g, (int)s->cfg.temporal_block)) {
            delete s;
            return fail(F2D_ERR_INVALID,
                        "F2D_JACOBI_STREAM needs cols %% 4 == 0 and temporal_block in {1,2,4,8}; use F2D_JACOBI_NAIVE");
        }
    } else {
        s->cfg.temporal_block = 1;
    }
    s->tune.chunk_rows = env_int("F2D_STREAM_CHUNK_ROWS", 0);
    s->tune.warps_per_cta = env_int("F2D_STREAM_WARPS_PER_CTA", 0);
    s->tune.rhs_in_smem = env_int("F2D_STREAM_RHS_SMEM", 0);

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
    for (int i = 0; i < 6; ++i) {
        if (cudaMalloc(&s->state[i], s->field_bytes) != cudaSuccess)
            return cleanup(fail(F2D_ERR_CUDA, "cudaMalloc(%zu) failed: %s", s->field_bytes, cudaGetErrorString(cudaGetLastError())));
        cudaMemsetAsync(s->state[i], 0, s->field_bytes, s->stream);
    }
    const int ntemps = 9;  // deepest point: velocity chain during project (2+2 diffuse, 2 advect, div, 2 p)
    for (int i = 0; i < ntemps; ++i) {
        float* p = nullptr;
        if (cudaMalloc(&p, s->field_bytes) != cudaSuccess)
            return cleanup(fail(F2D_ERR_CUDA, "cudaMalloc(%zu) failed: %s", s->field_bytes, cudaGetErrorString(cudaGetLastError())));
        cudaMemsetAsync(p, 0, s->field_bytes, s->stream);
        s->temps.push_back(p);
        s->temp_busy.push_back(0);
    }
    if (cudaMalloc(&s->oob_flag, sizeof(int)) != cudaSuccess)
        return cleanup(fail(F2D_ERR_CUDA, "cudaMalloc failed: %s", cudaGetErrorString(cudaGetLastError())));
    cudaMemsetAsync(s->oob_flag, 0, sizeof(int), s->stream);
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
    for (int i = 0; i < 6; ++i)
        if (s->state[i]) cudaFree(s->state[i]);
    for (float* p : s->temps) cudaFree(p);
    if (s->oob_flag) cudaFree(s->oob_flag);
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    if (s->own_stream && s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

#define F2D_NEED(s)                                              \
    do {                                                         \
        if (!(s)) return fail(F2D_ERR_INVALID, "solver is NULL"); \
        F2D_CUDA(cudaSetDevice((s)->device));                    \
    } while (0)

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
    if (oob) {
        cudaMemset(s->oob_flag, 0, sizeof(int));
        return fail(F2D_ERR_STATE, "density scatter left the slab: displacement exceeded the halo (CFL bound violated)");
    }
    return F2D_OK;
}

F2D_API int f2d_solve_host(f2d_solver* s, float* density, const float* density_source, float diffusion_rate, float* u, float* v,
                   const float* u_source, const float* v_source, float viscosity, float dt) {
    F2D_NEED(s);
    if (!density || !density_source || !u || !v || !u_source || !v_source) return fail(F2D_ERR_INVALID, "NULL grid pointer");
    // upload (gpu.cu:232-234 and the source uploads of :281) -> one step -> download (:255-257)
    F2D_TRY(s->h2d(s->state[F2D_FIELD_DENSITY], density));
    F2D_TRY(s->h2d(s->state[F2D_FIELD_U], u));
    F2D_TRY(s->h2d(s->state[F2D_FIELD_V], v));
    F2D_TRY(s->h2d(s->state[F2D_FIELD_DENSITY_SOURCE], density_source));
    F2D_TRY(s->h2d(s->state[F2D_FIELD_U_SOURCE], u_source));
    F2D_TRY(s->h2d(s->state[F2D_FIELD_V_SOURCE], v_source));
    F2D_TRY(s->step(diffusion_rate, viscosity, dt, 1));
    F2D_TRY(s->d2h(density, s->state[F2D_FIELD_DENSITY]));
    F2D_TRY(s->d2h(u, s->state[F2D_FIELD_U]));
    F2D_TRY(s->d2h(v, s->state[F2D_FIELD_V]));
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
    return F2D_OK;
}

F2D_API int f2d_stage_add_sources(f2d_solver* s, int field, float dt) {
    F2D_NEED(s);
    if (field < 0 || field > F2D_FIELD_V) return fail(F2D_ERR_INVALID, "bad field");
    AddSourceBatch ab;
    ab.n = 1;
    ab.f[0] = s->state[field];
    ab.s[0] = s->state[field + 3];
    launch_add_sources(s->g, ab, dt, s->stream);
    s->count();
    F2D_CUDA(cudaGetLastError());
    return F2D_OK;
}

F2D_API int f2d_stage_diffuse(f2d_solver* s, int field, int kind, float rate, float dt, uint32_t iters) {
    F2D_NEED(s);
    if (field < 0 || field > F2D_FIELD_V) return fail(F2D_ERR_INVALID, "bad field");
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
    launch_advect_velocity(s->g, u0, v0, u, v, s->dt0(dt), s->stream);
    s->count();
    s->release(u0);
    s->release(v0);
    F2D_CUDA(cudaGetLastError());
    return F2D_OK;
}

F2D_API int f2d_stage_project(f2d_solver* s, uint32_t iters) {
    F2D_NEED(s);
    float *u = s->state[F2D_FIELD_U], *v = s->state[F2D_FIELD_V];
    float *u0 = s->acquire(), *v0 = s->acquire();
    if (!u0 || !v0) return fail(F2D_ERR_STATE, "scratch pool exhausted");
    F2D_TRY(copy_back(s, u0, u));
    F2D_TRY(copy_back(s, v0, v));
    int rc = s->project(u0, v0, u, v, iters);
    s->release(u0);
    s->release(v0);
    return rc;
}

F2D_API int f2d_bench_jacobi(f2d_solver* s, int diffuse_like, uint32_t iters, uint32_t reps, float* elapsed_ms) {
    F2D_NEED(s);
    if (!elapsed_ms || reps == 0) return fail(F2D_ERR_INVALID, "bad arguments");
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
