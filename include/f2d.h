/*
 * f2d.h -- C ABI of the B200-native stable-fluids step (libf2d.so).
 *
 * This is the drop-in boundary for the ONE hot path of mworchel/fluid-2d:
 * fluid_solver::solve (src/fluid_solver.hpp:16-24) and the stages behind it
 * (src/fluid_solver_gpu.cu:222-404).  Plain pointers and sizes only; no C++,
 * CUDA or torch types cross this boundary.  Every entry point returns an int
 * status (F2D_OK == 0); f2d_last_error() describes the last failure of the
 * calling thread.  Nothing throws.  One f2d_solver may be used by one thread at
 * a time (the reference solver is not re-entrant either: member temp buffers,
 * src/fluid_solver_gpu.cuh:62-68).
 *
 * Arithmetic contract: by default the results follow fluid_solver_gpu (true Jacobi,
 * edges without corners, density `smooth`); iteration counts
 * are parameters instead of the literals 15/20 (src/fluid_solver_gpu.cu:238-252).
 * With f2d_config.semantics = F2D_SEM_CPU they follow fluid_solver_cpu bit for bit.
 *
 * The C++ adapter `fluid_solver_b200` (include/fluid_solver_b200.hpp) and the
 * Python host mirror (fluid-2d_b200/solver.py) are thin layers over these calls.
 */
#ifndef F2D_H_
#define F2D_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define F2D_ABI_VERSION 3

#if defined(__GNUC__)
#define F2D_API __attribute__((visibility("default")))
#else
#define F2D_API
#endif

/* status codes */
#define F2D_OK 0
#define F2D_ERR_INVALID 1   /* bad argument / configuration                  */
#define F2D_ERR_CUDA 2      /* a CUDA runtime call failed (see last_error)   */
#define F2D_ERR_NO_DEVICE 3 /* no usable CUDA device: there is NO CPU fallback */
#define F2D_ERR_STATE 4     /* call not valid in the solver's current state  */

/* fields of the solver state (replace the reference's m_*_buffer members,
 * src/fluid_solver_gpu.cuh:62-68) */
#define F2D_FIELD_DENSITY 0
#define F2D_FIELD_U 1 /* horizontal velocity */
#define F2D_FIELD_V 2 /* vertical velocity   */
#define F2D_FIELD_DENSITY_SOURCE 3
#define F2D_FIELD_U_SOURCE 4
#define F2D_FIELD_V_SOURCE 5
#define F2D_FIELD_PRESSURE 6   /* last pressure of the last project (debug/parity) */
#define F2D_FIELD_DIVERGENCE 7 /* last divergence of the last project             */
#define F2D_FIELD_COUNT 8

/* boundary kinds (src/fluid_solver_gpu.cu:11-54) */
#define F2D_BND_CONTINUOUS 0
#define F2D_BND_OPPOSITE_HORIZONTAL 1
#define F2D_BND_OPPOSITE_VERTICAL 2

/* relaxation kernels */
#define F2D_JACOBI_NAIVE 0  /* one sweep per launch, one thread per cell (bring-up / cross-check) */
#define F2D_JACOBI_STREAM 1 /* register-pipelined, temporally blocked row streaming (default)     */

/* how the diffuse sweep forms (x0 + a*sum4) / (1 + 4a)  (src/fluid_solver_gpu.cu:81-82) */
#define F2D_DIV_F64 0      /* (float)((double)num / (1.0 + 4.0*a)): the reference's own arithmetic */
#define F2D_DIV_F32_CORR 1 /* correctly rounded multiplication by the 48-bit reciprocal of 1+4a: FMUL + FMA
                              (default; differs from the fp64 divide only within ~2^-24 ulp of a rounding tie) */

/* which of the reference's two solvers the arithmetic follows (they are NOT numerically equivalent,
 * SURVEY.md Appendix B) */
#define F2D_SEM_GPU 0 /* fluid_solver_gpu (src/fluid_solver_gpu.cu): Jacobi, FMA where nvcc contracts, corners untouched,
                         density smooth, atomics in the scatter (default)                                            */
#define F2D_SEM_CPU 1 /* fluid_solver_cpu (src/fluid_solver_cpu.cpp): in-place lexicographic Gauss-Seidel, no FMA, float
                         divides, averaged corners, scatter summed in source order.  Bit-identical to that solver,
                         density included.  Single GPU; set diffuse_iters = project_iters = 20 and smooth = 0 to get
                         exactly fluid_solver_cpu::solve (cpp:15-30)                                                  */

typedef struct f2d_solver f2d_solver; /* opaque: owns device fields, stream, graphs */

typedef struct f2d_config {
    uint32_t struct_size;    /* sizeof(f2d_config), for ABI evolution                          */
    uint32_t rows, cols;     /* LOCAL field extent held by this solver (== global on one GPU)   */
    uint32_t diffuse_iters;  /* Kd; fluid_solver_gpu::solve uses 15 (gpu.cu:238,245-246)        */
    uint32_t project_iters;  /* Kp; fluid_solver_gpu::solve uses 20 (gpu.cu:247,252)            */
    uint32_t smooth;         /* 1 = density smooth after advect as gpu.cu:240 (default)         */
    uint32_t jacobi_mode;    /* F2D_JACOBI_*                                                    */
    uint32_t temporal_block; /* sweeps fused per pass of the PRESSURE solve (1,2,4,8); 0 = auto (4 up to ~1024^2, else 8) */
    uint32_t divide_mode;    /* F2D_DIV_*                                                       */
    uint32_t use_graph;      /* 1 = capture the step into a CUDA graph (default)                */
    int32_t device;          /* CUDA device ordinal; -1 = current device                        */
    /* row-slab decomposition (multi-GPU).  One GPU: global_rows = rows, row_offset = 0, halo = 0.
     * Local row i is global row row_offset + i; rows [0,halo) and [rows-halo,rows) are halo rows
     * owned by the neighbour slabs unless they touch the global top/bottom edge.               */
    uint32_t global_rows;
    uint32_t row_offset;
    uint32_t halo;
    uint32_t temporal_block_diffuse; /* same for the diffuse solve; 0 = auto (temporal_block if set, else as above) */
    uint32_t semantics;      /* F2D_SEM_*; jacobi_mode, temporal_block* and divide_mode only apply to F2D_SEM_GPU */
    void* stream; /* cudaStream_t to run on; NULL = the solver creates its own                  */
} f2d_config;

/* Fill *cfg with defaults for a rows x cols single-GPU solver (Kd=15, Kp=20, smooth=1: exactly
 * fluid_solver_gpu::solve). */
F2D_API int f2d_config_default(f2d_config* cfg, uint32_t rows, uint32_t cols);

/* fluid_solver_gpu::fluid_solver_gpu(rows, cols)  (src/fluid_solver_gpu.cu:209-218): allocate all
 * device fields once.  Fails with F2D_ERR_NO_DEVICE when no GPU is present. */
F2D_API int f2d_create(const f2d_config* cfg, f2d_solver** out);
F2D_API void f2d_destroy(f2d_solver* s);

/* ---- the reference interface ------------------------------------------------------------
 * fluid_solver::solve (src/fluid_solver.hpp:16-24; GPU: src/fluid_solver_gpu.cu:222-258):
 * host arrays (row-major, pitch == cols, like grid<float>::data()), density/u/v updated in
 * place, sources const.  Upload -> one device step -> download; blocking.  On one GPU with fields of >= 1 MiB the
 * three phases overlap: uploads in the order u, su, v, sv, sd, d; the step cut into four parts that start as their
 * inputs land; the velocity downloaded while the density is still being solved.  The
 * device-resident state afterwards equals the host grids. */
F2D_API int f2d_solve_host(f2d_solver* s, float* density, const float* density_source, float diffusion_rate,
                   float* u, float* v, const float* u_source, const float* v_source, float viscosity,
                   float dt);

/* Optional: page-lock caller-owned host memory so that f2d_solve_host / f2d_upload / f2d_download move it by DMA
 * (grid<float> storage is pageable; the reference pays the driver's staged copy on every call,
 * src/utilities.hpp:57-67).  LIFETIME CONTRACT: [host, host + bytes) must stay allocated and unmoved until
 * f2d_unpin_host(host) or f2d_destroy.  The library never registers caller memory on its own: freeing a registered
 * range leaves stale pinned pages behind that a later allocation at the same address would alias.  Memory that is
 * already page-locked (cudaHostAlloc, torch pin_memory) needs no call.  f2d_unpin_host waits for the solver's
 * streams first.  (Setting the environment variable F2D_HOST_REGISTER=1 restores round-1 behaviour -- every grid
 * of >= 1 MiB passed to f2d_solve_host is registered until f2d_destroy -- for callers that accept this contract
 * for all their grids.) */
F2D_API int f2d_pin_host(f2d_solver* s, const void* host, size_t bytes);
F2D_API int f2d_unpin_host(f2d_solver* s, const void* host);

/* ---- device-resident extension (the per-call PCIe copies are the interface's tax) -------- */
/* copy(m_*_buffer, grid)  (src/utilities.hpp:57-59).  NULL pointers are skipped. */
F2D_API int f2d_upload(f2d_solver* s, const float* density, const float* u, const float* v);
F2D_API int f2d_set_sources(f2d_solver* s, const float* density_source, const float* u_source,
                    const float* v_source);
/* copy(grid, m_*_buffer)  (src/utilities.hpp:65-67) */
F2D_API int f2d_download(f2d_solver* s, float* density, float* u, float* v);
/* generic single-field transfer; host pitch == cols */
F2D_API int f2d_upload_field(f2d_solver* s, int field, const float* host);
F2D_API int f2d_download_field(f2d_solver* s, int field, float* host);
/* simulation::update zeroes the sources after each solve (src/simulation.cpp:62-64) */
F2D_API int f2d_clear_sources(f2d_solver* s);

/* `nsteps` solve() steps on the device-resident state with constant sources; asynchronous on the
 * solver's stream. */
F2D_API int f2d_step(f2d_solver* s, float diffusion_rate, float viscosity, float dt, uint32_t nsteps);
/* same, bracketed by CUDA events on the solver's stream; blocks; *elapsed_ms = device time */
F2D_API int f2d_step_timed(f2d_solver* s, float diffusion_rate, float viscosity, float dt, uint32_t nsteps,
                   float* elapsed_ms);
F2D_API int f2d_sync(f2d_solver* s);

/* ---- single stages on the device-resident state (parity tests; each mirrors one private
 * method of fluid_solver_gpu, src/fluid_solver_gpu.cuh:27-56) ------------------------------- */
F2D_API int f2d_stage_set_bnd(f2d_solver* s, int field, int kind);           /* gpu.cu:260-276 */
F2D_API int f2d_stage_add_sources(f2d_solver* s, int field, float dt);       /* gpu.cu:278-288; field d/u/v, its own source */
F2D_API int f2d_stage_diffuse(f2d_solver* s, int field, int kind, float rate, float dt, uint32_t iters); /* gpu.cu:290-312 */
F2D_API int f2d_stage_smooth(f2d_solver* s);                                  /* gpu.cu:314-323 (density) */
F2D_API int f2d_stage_advect_density(f2d_solver* s, float dt);                /* gpu.cu:325-356, trace=true, by current u,v */
F2D_API int f2d_stage_advect_velocity(f2d_solver* s, float dt);               /* gpu.cu:248-251: U0,V0 <- u,v; advect u; advect v */
F2D_API int f2d_stage_project(f2d_solver* s, uint32_t iters);                 /* gpu.cu:358-404 */

/* ---- measurement helpers -------------------------------------------------------------------
 * Run `reps` pressure-relaxation solves of `iters` Jacobi sweeps each on scratch fields with the
 * configured kernel and return the device time of all of them (CUDA events on the solver stream).
 * Used by bench.py for the "Jacobi HBM GB/s vs peak" roofline line. */
F2D_API int f2d_bench_jacobi(f2d_solver* s, int diffuse_like, uint32_t iters, uint32_t reps, float* elapsed_ms);
/* number of kernel launches (graph kernel nodes included) issued by this solver so far */
F2D_API int f2d_launch_count(const f2d_solver* s, uint64_t* launches);

/* ---- multi-GPU row slabs (one process per GPU) ------------------------------------------------
 * New work: the reference is single-GPU.  Each rank holds one slab of rows
 * (f2d_config.row_offset / global_rows / halo); neighbouring slabs exchange `halo` rows over
 * NCCL send/recv (NVLink) whenever a stencil stage would otherwise invalidate owned rows: once per
 * halo/temporal_block relaxation passes, around the advection (radius cfl_cells + 1), and -- in the
 * reverse direction, with an add -- after the density scatter.  The exchanges run on the solver's
 * stream and are part of the step's CUDA graph.
 *   f2d_comm_unique_id : rank 0 creates the 128-byte NCCL id; the host side broadcasts it.
 *   f2d_comm_init      : collective; rank r must hold the r-th slab from the top.  cfl_cells is
 *                        the caller's bound on the advection displacement per step (cells);
 *                        f2d_sync reports F2D_ERR_STATE if a step exceeded it. */
F2D_API int f2d_comm_unique_id(char* id128);
F2D_API int f2d_comm_init(f2d_solver* s, const char* id128, int rank, int nranks, int cfl_cells);
F2D_API int f2d_comm_stats(const f2d_solver* s, uint64_t* exchanges);
/* payload bytes this rank has pushed to its upper / lower neighbour so far (halo rows, both directions of the
 * scatter's reverse exchange included); 0 at a global edge */
F2D_API int f2d_comm_bytes(const f2d_solver* s, uint64_t* to_up, uint64_t* to_down);
/* The same slabs with the halo transport replaced by direct NVLink peer stores (f2d_p2p.cu): one kernel
 * per exchange pushes the rows into the neighbour's halo and handshakes through epoch flags in device
 * memory; no NCCL call on the data path.  Every rank exports one 64-byte CUDA IPC handle of its arena plus
 * four words of geometry; the host side gathers them and passes each rank its neighbours' (NULL at the
 * global edges).  Use either f2d_comm_init or f2d_p2p_connect, not both. */
F2D_API int f2d_p2p_export(f2d_solver* s, unsigned char* handle64, uint64_t* info4);
F2D_API int f2d_p2p_connect(f2d_solver* s, int rank, int nranks, const unsigned char* up_handle64, const uint64_t* up_info4,
                            const unsigned char* down_handle64, const uint64_t* down_info4, int cfl_cells);

/* ---- headless renderers (next-row f3; read the device-resident fields, no window system) -------
 * f2d_render_density_rgba : grid_to_image_kernel (src/density_grid_renderer.cu:10-29): rows*cols RGBA8
 *                           pixels, channel = uint8(clamp(multiplier * density, 0, 255)), A = 255.
 * f2d_render_velocity_lines: velocity_to_lines_kernel (src/velocity_grid_renderer.cu:8-44): rows*cols
 *                           segments as 4 floats (start.x, start.y, end.x, end.y); every 8th row/column
 *                           the end follows the velocity (250000 * vel / sqrtf(rows*cols)). */
F2D_API int f2d_render_density_rgba(f2d_solver* s, float mult_r, float mult_g, float mult_b, unsigned char* host_rgba);
F2D_API int f2d_render_velocity_lines(f2d_solver* s, float horizontal_scale, float vertical_scale, float* host_lines);

/* ---- interop ------------------------------------------------------------------------------- */
/* device pointer + pitch (in floats) of a state field, for zero-copy views (torch, NCCL halos) */
F2D_API int f2d_field_ptr(f2d_solver* s, int field, void** device_ptr, size_t* pitch_elems);
F2D_API int f2d_get_config(const f2d_solver* s, f2d_config* out);
/* the stream the solver launches on (cudaStream_t as void*) */
F2D_API int f2d_get_stream(const f2d_solver* s, void** stream);

F2D_API const char* f2d_last_error(void);
F2D_API int f2d_abi_version(void);
F2D_API int f2d_device_count(void);

#ifdef __cplusplus
}
#endif
#endif /* F2D_H_ */
