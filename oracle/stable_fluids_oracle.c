/*
 * stable_fluids_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, single-threaded CPU restatement of the stable-fluids time step of
 * mworchel/fluid-2d, used as the parity checker for the CUDA path.  It is NOT
 * part of the product: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product
 * (fluid-2d_b200/csrc) never links or calls anything in this directory.
 *
 * Two arithmetic "semantics" are restated, because the reference ships two
 * solvers that are NOT numerically equivalent:
 *
 *   SFO_SEM_GPU  follows src/fluid_solver_gpu.cu (kernels :11-206, step :222-258)
 *                as nvcc 12.9 compiles it for sm_100a with default flags:
 *                true Jacobi relaxations, FMA contraction exactly where the
 *                SASS shows FFMA, fp64 divide in diffuse, edges without
 *                corners, the extra density `smooth`, FTZ on the scatter adds.
 *                THIS is the parity target of the CUDA path.
 *   SFO_SEM_CPU  follows src/fluid_solver_cpu.cpp (:6-215) as g++ -O2 compiles
 *                it for generic x86-64: in-place Gauss-Seidel, no FMA, float
 *                divide, averaged corners, no smooth.  It exists so that this
 *                file can be pinned bit-for-bit against the unmodified
 *                reference CPU solver (oracle/_ref/libref_cpu.so) and the
 *                anchors of SURVEY.md Appendix D.
 *
 * Pinning status: SFO_SEM_CPU is pinned bitwise by tests/test_oracle_pin.py
 * (reference CPU solver compiled from /root/reference + the FNV anchors).
 * SFO_SEM_GPU is pinned against fixtures produced by the unmodified
 * fluid_solver_gpu run on a B200 (tests/golden/refgpu_*.npz, generator
 * oracle/ref_gpu_driver.cu).  The reference has no tests or golden vectors of
 * its own.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math (see oracle/Makefile); fmaf()
 * is used explicitly wherever the reference SASS contracts.
 *
 * Layout: row-major, element (i=row, j=col) at f[i*cols + j], like grid<T>
 * (src/grid.hpp:35-44).
 */
#include <math.h>
#include <float.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SFO_SEM_GPU 0
#define SFO_SEM_CPU 1

#define SFO_BND_CONTINUOUS 0          /* density, pressure, divergence */
#define SFO_BND_OPPOSITE_HORIZONTAL 1 /* u: left/right columns negated  */
#define SFO_BND_OPPOSITE_VERTICAL 2   /* v: top/bottom rows negated     */

#define AT(f, i, j) ((f)[(size_t)(i) * cols + (size_t)(j)])

static float *dup_field(const float *f, size_t rows, size_t cols) {
    float *c = (float *)malloc(sizeof(float) * rows * cols);
    memcpy(c, f, sizeof(float) * rows * cols);
    return c;
}

/* ------------------------------------------------------------------ set_bnd
 * GPU: src/fluid_solver_gpu.cu:11-54 -- n = 1..N-2 only, corners never written.
 * CPU: src/fluid_solver_cpu.cpp:33-83 -- all n (so the first loop also writes
 *      the corner rows' edge cells), second loop overwrites, then corners are
 *      the mean of their two edge neighbours.
 * For the CPU the loop ORDER differs for opposite_vertical (rows loop first,
 * cpp:69-77) which matters only for the corner-adjacent cells that the final
 * corner average then overwrites; it is restated literally anyway.          */
void sfo_set_bnd(float *f, size_t rows, size_t cols, int kind, int sem) {
    if (sem == SFO_SEM_GPU) {
        float sc = (kind == SFO_BND_OPPOSITE_HORIZONTAL) ? -1.f : 1.f; /* sign on columns */
        float sr = (kind == SFO_BND_OPPOSITE_VERTICAL) ? -1.f : 1.f;   /* sign on rows    */
        for (size_t i = 1; i + 1 < rows; ++i) {
            AT(f, i, 0) = (sc < 0) ? -AT(f, i, 1) : AT(f, i, 1);
            AT(f, i, cols - 1) = (sc < 0) ? -AT(f, i, cols - 2) : AT(f, i, cols - 2);
        }
        for (size_t j = 1; j + 1 < cols; ++j) {
            AT(f, 0, j) = (sr < 0) ? -AT(f, 1, j) : AT(f, 1, j);
            AT(f, rows - 1, j) = (sr < 0) ? -AT(f, rows - 2, j) : AT(f, rows - 2, j);
        }
        return;
    }
    /* CPU semantics */
    if (kind == SFO_BND_OPPOSITE_VERTICAL) {
        for (size_t j = 0; j < cols; ++j) {
            AT(f, 0, j) = -AT(f, 1, j);
            AT(f, rows - 1, j) = -AT(f, rows - 2, j);
        }
        for (size_t i = 0; i < rows; ++i) {
            AT(f, i, 0) = AT(f, i, 1);
            AT(f, i, cols - 1) = AT(f, i, cols - 2);
        }
    } else {
        int neg = (kind == SFO_BND_OPPOSITE_HORIZONTAL);
        for (size_t i = 0; i < rows; ++i) {
            AT(f, i, 0) = neg ? -AT(f, i, 1) : AT(f, i, 1);
            AT(f, i, cols - 1) = neg ? -AT(f, i, cols - 2) : AT(f, i, cols - 2);
        }
        for (size_t j = 0; j < cols; ++j) {
            AT(f, 0, j) = AT(f, 1, j);
            AT(f, rows - 1, j) = AT(f, rows - 2, j);
        }
    }
    AT(f, 0, 0) = 0.5f * (AT(f, 0, 1) + AT(f, 1, 0));
    AT(f, 0, cols - 1) = 0.5f * (AT(f, 0, cols - 2) + AT(f, 1, cols - 1));
    AT(f, rows - 1, 0) = 0.5f * (AT(f, rows - 1, 1) + AT(f, rows - 2, 0));
    AT(f, rows - 1, cols - 1) = 0.5f * (AT(f, rows - 1, cols - 2) + AT(f, rows - 2, cols - 1));
}

/* -------------------------------------------------------------- add_sources
 * GPU: src/fluid_solver_gpu.cu:56-67 (one FFMA(dt, s, f)); CPU: cpp:85-93
 * (product rounded first).  Interior only, no boundary pass afterwards.      */
void sfo_add_sources(float *f, const float *s, size_t rows, size_t cols, float dt, int sem) {
    for (size_t i = 1; i + 1 < rows; ++i)
        for (size_t j = 1; j + 1 < cols; ++j) {
            if (sem == SFO_SEM_GPU)
                AT(f, i, j) = fmaf(dt, AT(s, i, j), AT(f, i, j));
            else {
                float t = dt * AT(s, i, j);
                AT(f, i, j) = AT(f, i, j) + t;
            }
        }
}

/* the relaxation coefficient a = dt * float(rows*cols) * rate, fp32, left to right
 * (gpu.cu:79, cpu.cpp:100) */
float sfo_diffuse_coeff(size_t rows, size_t cols, float rate, float dt) {
    float t = dt * (float)(rows * cols);
    return t * rate;
}

/* ------------------------------------------------------------------ diffuse
 * GPU: host loop src/fluid_solver_gpu.cu:290-312, kernel :69-85.
 *      x0 <- f (whole field); K times { prev <- f; interior:
 *      num = FFMA(a, ((W+E)+N)+S, x0); f = (float)((double)num / (1.0 + 4.0*a)); set_bnd }.
 * CPU: src/fluid_solver_cpu.cpp:95-114, Gauss-Seidel in place,
 *      (x0 + a*(((N+S)+W)+E)) / (1.f + 4.f*a).                                 */
void sfo_diffuse(float *f, size_t rows, size_t cols, int kind, float rate, float dt,
                 unsigned iters, int sem) {
    float a = sfo_diffuse_coeff(rows, cols, rate, dt);
    float *x0 = dup_field(f, rows, cols);
    if (sem == SFO_SEM_GPU) {
        double c = 1.0 + 4.0 * (double)a;
        float *prev = (float *)malloc(sizeof(float) * rows * cols);
        for (unsigned k = 0; k < iters; ++k) {
            memcpy(prev, f, sizeof(float) * rows * cols);
            for (size_t i = 1; i + 1 < rows; ++i)
                for (size_t j = 1; j + 1 < cols; ++j) {
                    float s = AT(prev, i, j - 1) + AT(prev, i, j + 1);
                    s = s + AT(prev, i - 1, j);
                    s = s + AT(prev, i + 1, j);
                    float num = fmaf(a, s, AT(x0, i, j));
                    AT(f, i, j) = (float)((double)num / c);
                }
            sfo_set_bnd(f, rows, cols, kind, sem);
        }
        free(prev);
    } else {
        float c = 1.f + 4.f * a;
        for (unsigned k = 0; k < iters; ++k) {
            for (size_t i = 1; i + 1 < rows; ++i)
                for (size_t j = 1; j + 1 < cols; ++j) {
                    float s = AT(f, i - 1, j) + AT(f, i + 1, j);
                    s = s + AT(f, i, j - 1);
                    s = s + AT(f, i, j + 1);
                    float t = a * s;
                    float num = AT(x0, i, j) + t;
                    AT(f, i, j) = num / c;
                }
            sfo_set_bnd(f, rows, cols, kind, sem);
        }
    }
    free(x0);
}

/* dt0 = sqrt(rows*cols) * dt.  GPU: double sqrt(size_t) * float -> float
 * (gpu.cu:334); CPU: sqrtf(float(rows*cols)) * dt (cpp:126).                   */
float sfo_dt0(size_t rows, size_t cols, float dt, int sem) {
    if (sem == SFO_SEM_GPU) return (float)(sqrt((double)(rows * cols)) * (double)dt);
    return sqrtf((float)(rows * cols)) * dt;
}

/* ------------------------------------------------------------ advect, gather
 * trace=false path.  GPU: src/fluid_solver_gpu.cu:99-129 (+ host :325-356);
 * CPU: src/fluid_solver_cpu.cpp:153-174.  Back-trace, clamp to [1.5, N-1.5],
 * bilinear gather from a copy of f, then set_bnd.  uu/vv must not alias f.     */
void sfo_advect_gather(float *f, const float *uu, const float *vv, size_t rows, size_t cols,
                       int kind, float dt, int sem) {
    float *src = dup_field(f, rows, cols);
    float dt0 = sfo_dt0(rows, cols, dt, sem);
    float xmax = (float)cols - 1.5f, ymax = (float)rows - 1.5f;
    for (size_t i = 1; i + 1 < rows; ++i)
        for (size_t j = 1; j + 1 < cols; ++j) {
            float x, y;
            if (sem == SFO_SEM_GPU) {
                x = fmaf(-AT(uu, i, j), dt0, (float)j);
                y = fmaf(-AT(vv, i, j), dt0, (float)i);
            } else {
                float tx = dt0 * AT(uu, i, j), ty = dt0 * AT(vv, i, j);
                x = (float)j - tx;
                y = (float)i - ty;
            }
            x = fmaxf(1.5f, fminf(xmax, x));
            y = fmaxf(1.5f, fminf(ymax, y));
            size_t j0 = (size_t)x, i0 = (size_t)y, j1 = j0 + 1, i1 = i0 + 1;
            float s0 = x - (float)j0, s1 = 1.f - s0, s2 = y - (float)i0, s3 = 1.f - s2;
            float a00 = AT(src, i0, j0), a01 = AT(src, i0, j1);
            float a10 = AT(src, i1, j0), a11 = AT(src, i1, j1);
            if (sem == SFO_SEM_GPU) {
                /* SASS: FMUL,FMUL,FFMA,FMUL,FFMA,FFMA */
                float top = fmaf(s0, a01, s1 * a00);
                float bot = fmaf(s0, a11, s1 * a10);
                float t = s2 * bot;
                AT(f, i, j) = fmaf(s3, top, t);
            } else {
                float p0 = s1 * a00, p1 = s0 * a01, p2 = s1 * a10, p3 = s0 * a11;
                float top = p0 + p1, bot = p2 + p3;
                float t0 = s3 * top, t1 = s2 * bot;
                AT(f, i, j) = t0 + t1;
            }
        }
    free(src);
    sfo_set_bnd(f, rows, cols, kind, sem);
}

static inline float ftz(float x) { return (fabsf(x) < FLT_MIN) ? copysignf(0.f, x) : x; }

/* ----------------------------------------------------------- advect, scatter
 * trace=true path (density).  GPU: src/fluid_solver_gpu.cu:131-162 (+ :336-345):
 * f <- 0 (whole buffer), forward trace, skip when outside [0.5, N-1.5], four
 * atomicAdd (RED.ADD.F32.FTZ) per source cell in hardware order -- restated here
 * in lexicographic order, so GPU results match only to summation-order rounding.
 * CPU: src/fluid_solver_cpu.cpp:127-152 (lexicographic, no FTZ).               */
void sfo_advect_scatter(float *f, const float *uu, const float *vv, size_t rows, size_t cols,
                        int kind, float dt, int sem) {
    float *src = dup_field(f, rows, cols);
    float dt0 = sfo_dt0(rows, cols, dt, sem);
    float xmax = (float)cols - 1.5f, ymax = (float)rows - 1.5f;
    memset(f, 0, sizeof(float) * rows * cols);
    for (size_t i = 1; i + 1 < rows; ++i)
        for (size_t j = 1; j + 1 < cols; ++j) {
            float x, y;
            if (sem == SFO_SEM_GPU) {
                x = fmaf(AT(uu, i, j), dt0, (float)j);
                y = fmaf(AT(vv, i, j), dt0, (float)i);
            } else {
                float tx = dt0 * AT(uu, i, j), ty = dt0 * AT(vv, i, j);
                x = (float)j + tx;
                y = (float)i + ty;
            }
            if (x < 0.5f || x > xmax || y < 0.5f || y > ymax) continue;
            size_t j0 = (size_t)x, i0 = (size_t)y, j1 = j0 + 1, i1 = i0 + 1;
            float s0 = x - (float)j0, s1 = 1.f - s0, s2 = y - (float)i0, s3 = 1.f - s2;
            float v = AT(src, i, j);
            float w00 = s1 * s3, w10 = s1 * s2, w01 = s0 * s3, w11 = s0 * s2;
            float c00 = w00 * v, c10 = w10 * v, c01 = w01 * v, c11 = w11 * v;
            if (sem == SFO_SEM_GPU) {
                AT(f, i0, j0) = ftz(ftz(AT(f, i0, j0)) + ftz(c00));
                AT(f, i1, j0) = ftz(ftz(AT(f, i1, j0)) + ftz(c10));
                AT(f, i0, j1) = ftz(ftz(AT(f, i0, j1)) + ftz(c01));
                AT(f, i1, j1) = ftz(ftz(AT(f, i1, j1)) + ftz(c11));
            } else {
                AT(f, i0, j0) += c00;
                AT(f, i1, j0) += c10;
                AT(f, i0, j1) += c01;
                AT(f, i1, j1) += c11;
            }
        }
    free(src);
    sfo_set_bnd(f, rows, cols, kind, sem);
}

/* ------------------------------------------------------------------- smooth
 * GPU only: src/fluid_solver_gpu.cu:314-323, kernel :87-97.
 * f = 0.2f * ((((c + W) + E) + N) + S) on the interior; NO set_bnd afterwards. */
void sfo_smooth(float *f, size_t rows, size_t cols) {
    float *src = dup_field(f, rows, cols);
    for (size_t i = 1; i + 1 < rows; ++i)
        for (size_t j = 1; j + 1 < cols; ++j) {
            float s = AT(src, i, j) + AT(src, i, j - 1);
            s = s + AT(src, i, j + 1);
            s = s + AT(src, i - 1, j);
            s = s + AT(src, i + 1, j);
            AT(f, i, j) = 0.2f * s;
        }
    free(src);
}

/* ------------------------------------------------------------------ project
 * GPU: src/fluid_solver_gpu.cu:358-404, kernels :164-206.  CPU: cpp:179-215.
 * div = (-0.5f*h) * (((uE - uW) + vS) - vN), set_bnd_cont(div); p = 0;
 * K relaxations p = ((((div + pE) + pW) + pS) + pN) * 0.25f with set_bnd_cont(p)
 * (Jacobi on the GPU, Gauss-Seidel on the CPU); then
 * u -= (0.5f*(pE - pW)) / h, v -= (0.5f*(pS - pN)) / h and the two velocity
 * boundary passes.  If p_out/div_out are non-NULL the final p/div are copied out. */
void sfo_project(float *u, float *v, size_t rows, size_t cols, unsigned iters, int sem,
                 float *p_out, float *div_out) {
    size_t n = rows * cols;
    float h = 1.0f / sqrtf((float)(rows * cols));
    float *dv = (float *)calloc(n, sizeof(float));
    float *p = (float *)calloc(n, sizeof(float));
    float mh = -0.5f * h;
    for (size_t i = 1; i + 1 < rows; ++i)
        for (size_t j = 1; j + 1 < cols; ++j) {
            float s = AT(u, i, j + 1) - AT(u, i, j - 1);
            s = s + AT(v, i + 1, j);
            s = s - AT(v, i - 1, j);
            AT(dv, i, j) = mh * s;
        }
    sfo_set_bnd(dv, rows, cols, SFO_BND_CONTINUOUS, sem);
    if (sem == SFO_SEM_GPU) {
        float *pp = (float *)malloc(sizeof(float) * n);
        for (unsigned k = 0; k < iters; ++k) {
            memcpy(pp, p, sizeof(float) * n);
            for (size_t i = 1; i + 1 < rows; ++i)
                for (size_t j = 1; j + 1 < cols; ++j) {
                    float s = AT(dv, i, j) + AT(pp, i, j + 1);
                    s = s + AT(pp, i, j - 1);
                    s = s + AT(pp, i + 1, j);
                    s = s + AT(pp, i - 1, j);
                    AT(p, i, j) = s * 0.25f;
                }
            sfo_set_bnd(p, rows, cols, SFO_BND_CONTINUOUS, sem);
        }
        free(pp);
    } else {
        for (unsigned k = 0; k < iters; ++k) {
            for (size_t i = 1; i + 1 < rows; ++i)
                for (size_t j = 1; j + 1 < cols; ++j) {
                    float s = AT(dv, i, j) + AT(p, i, j + 1);
                    s = s + AT(p, i, j - 1);
                    s = s + AT(p, i + 1, j);
                    s = s + AT(p, i - 1, j);
                    AT(p, i, j) = s / 4.0f;
                }
            sfo_set_bnd(p, rows, cols, SFO_BND_CONTINUOUS, sem);
        }
    }
    for (size_t i = 1; i + 1 < rows; ++i)
        for (size_t j = 1; j + 1 < cols; ++j) {
            float gx = 0.5f * (AT(p, i, j + 1) - AT(p, i, j - 1));
            float gy = 0.5f * (AT(p, i + 1, j) - AT(p, i - 1, j));
            gx = gx / h;
            gy = gy / h;
            AT(u, i, j) = AT(u, i, j) - gx;
            AT(v, i, j) = AT(v, i, j) - gy;
        }
    sfo_set_bnd(u, rows, cols, SFO_BND_OPPOSITE_HORIZONTAL, sem);
    sfo_set_bnd(v, rows, cols, SFO_BND_OPPOSITE_VERTICAL, sem);
    if (p_out) memcpy(p_out, p, sizeof(float) * n);
    if (div_out) memcpy(div_out, dv, sizeof(float) * n);
    free(dv);
    free(p);
}

/* --------------------------------------------------------------------- step
 * One full solve() with free iteration counts.
 * GPU order: src/fluid_solver_gpu.cu:232-257 (Kd=15, Kp=20, smooth on in the
 * reference); CPU order: src/fluid_solver_cpu.cpp:15-30 (K=20, no smooth).
 * The density chain uses the PRE-step u,v.                                      */
void sfo_step(float *d, const float *sd, float diffusion_rate, float *u, float *v,
              const float *su, const float *sv, float viscosity, float dt, size_t rows,
              size_t cols, unsigned kd, unsigned kp, int do_smooth, int sem) {
    sfo_add_sources(d, sd, rows, cols, dt, sem);
    sfo_diffuse(d, rows, cols, SFO_BND_CONTINUOUS, diffusion_rate, dt, kd, sem);
    sfo_advect_scatter(d, u, v, rows, cols, SFO_BND_CONTINUOUS, dt, sem);
    if (do_smooth) sfo_smooth(d, rows, cols);

    sfo_add_sources(u, su, rows, cols, dt, sem);
    sfo_add_sources(v, sv, rows, cols, dt, sem);
    sfo_diffuse(u, rows, cols, SFO_BND_OPPOSITE_HORIZONTAL, viscosity, dt, kd, sem);
    sfo_diffuse(v, rows, cols, SFO_BND_OPPOSITE_VERTICAL, viscosity, dt, kd, sem);
    sfo_project(u, v, rows, cols, kp, sem, NULL, NULL);
    float *u0 = dup_field(u, rows, cols), *v0 = dup_field(v, rows, cols);
    sfo_advect_gather(u, u0, v0, rows, cols, SFO_BND_OPPOSITE_HORIZONTAL, dt, sem);
    sfo_advect_gather(v, u0, v0, rows, cols, SFO_BND_OPPOSITE_VERTICAL, dt, sem);
    free(u0);
    free(v0);
    sfo_project(u, v, rows, cols, kp, sem, NULL, NULL);
}

void sfo_steps(float *d, const float *sd, float diffusion_rate, float *u, float *v,
               const float *su, const float *sv, float viscosity, float dt, size_t rows,
               size_t cols, unsigned kd, unsigned kp, int do_smooth, int sem, unsigned nsteps) {
    for (unsigned s = 0; s < nsteps; ++s)
        sfo_step(d, sd, diffusion_rate, u, v, su, sv, viscosity, dt, rows, cols, kd, kp,
                 do_smooth, sem);
}

/* ------------------------------------------------- canonical synthetic fields
 * SURVEY.md section 8(d) / BASELINE.md section 4: evaluated in double with libm,
 * then rounded to fp32.  Rows [row_begin,row_end) are written so that callers
 * can fill large grids from several threads; pointers address the FULL fields. */
static uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

void sfo_canonical_fields(size_t n, size_t row_begin, size_t row_end, float *d, float *u,
                          float *v, float *sd, float *su, float *sv) {
    const double pi = 3.14159265358979323846;
    const uint64_t seed = 0x2D5F1DULL;
    const double A = 4.0 / ((double)n * 0.02), m = 2.0;
    size_t cols = n;
    for (size_t i = row_begin; i < row_end; ++i)
        for (size_t j = 0; j < n; ++j) {
            double x = ((double)j + 0.5) / (double)n, y = ((double)i + 0.5) / (double)n;
            double U = (double)(splitmix64(seed + (uint64_t)i * n + j) >> 40) * (1.0 / 16777216.0);
            if (u) AT(u, i, j) = (float)(A * sin(2.0 * pi * m * x) * cos(2.0 * pi * m * y));
            if (v) AT(v, i, j) = (float)(-A * cos(2.0 * pi * m * x) * sin(2.0 * pi * m * y));
            if (d)
                AT(d, i, j) = (float)(exp(-((x - 0.5) * (x - 0.5) + (y - 0.5) * (y - 0.5)) /
                                          (2.0 * 0.1 * 0.1)) +
                                      0.05 * U);
            int in_disc = ((x - 0.5) * (x - 0.5) + (y - 0.25) * (y - 0.25)) < 0.05 * 0.05;
            if (sd) AT(sd, i, j) = in_disc ? 1.0f : 0.0f;
            if (sv) AT(sv, i, j) = in_disc ? (float)A : 0.0f;
            if (su) AT(su, i, j) = 0.0f;
        }
}

/* FNV-1a-64 over raw bytes (anchor hashes of SURVEY.md Appendix D). */
uint64_t sfo_fnv1a64(const void *data, size_t nbytes) {
    const unsigned char *p = (const unsigned char *)data;
    uint64_t h = 14695981039346656037ULL;
    for (size_t k = 0; k < nbytes; ++k) {
        h ^= p[k];
        h *= 1099511628211ULL;
    }
    return h;
}
