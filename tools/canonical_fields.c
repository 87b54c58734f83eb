/* canonical_fields.c -- the synthetic benchmark inputs of SURVEY.md section 8(d) (seed 0x2D5F1D).
 *
 * Workload generator for bench.py / bench_multi.py / tools: neutral code, neither product nor oracle
 * (the product arm of the bench must not load anything under oracle/).  Every value is evaluated in
 * double and rounded to fp32 once, so any host reproduces the same bits:
 *   x = (j+0.5)/N, y = (i+0.5)/N, A = 4/(N*0.02) (largest displacement 4 cells per step), m = 2
 *   u0 =  A sin(2 pi m x) cos(2 pi m y),  v0 = -A cos(2 pi m x) sin(2 pi m y)       (Taylor-Green)
 *   d0 = exp(-((x-1/2)^2 + (y-1/2)^2) / (2*0.1^2)) + 0.05 * U(i,j),  U = (splitmix64(seed + i*N + j) >> 40) * 2^-24
 *   sources on the disc (x-1/2)^2 + (y-1/4)^2 < 0.05^2: sd = 1, sv = A, su = 0.
 * Rows [row_begin, row_end) of the N x N grid are written at their GLOBAL position (pointer + i*N + j); a NULL
 * pointer skips that field.  Thread-safe (callers split the row range).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

static uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

void f2d_canonical_fields(size_t n, size_t row_begin, size_t row_end, float *d, float *u, float *v, float *sd,
                          float *su, float *sv) {
    const double two_pi_m = 2.0 * 3.14159265358979323846 * 2.0;
    const double amp = 4.0 / ((double)n * 0.02);
    for (size_t i = row_begin; i < row_end; ++i) {
        const double y = ((double)i + 0.5) / (double)n;
        const double sy = sin(two_pi_m * y), cy = cos(two_pi_m * y);
        const double dy2 = (y - 0.5) * (y - 0.5), sy2 = (y - 0.25) * (y - 0.25);
        for (size_t j = 0; j < n; ++j) {
            const size_t o = i * n + j;
            const double x = ((double)j + 0.5) / (double)n;
            const double dx2 = (x - 0.5) * (x - 0.5);
            if (u) u[o] = (float)(amp * sin(two_pi_m * x) * cy);
            if (v) v[o] = (float)(-amp * cos(two_pi_m * x) * sy);
            if (d) {
                const double noise = (double)(mix64(0x2D5F1DULL + (uint64_t)i * n + j) >> 40) * (1.0 / 16777216.0);
                d[o] = (float)(exp(-(dx2 + dy2) / (2.0 * 0.1 * 0.1)) + 0.05 * noise);
            }
            const int disc = (dx2 + sy2) < 0.05 * 0.05;
            if (sd) sd[o] = disc ? 1.0f : 0.0f;
            if (sv) sv[o] = disc ? (float)amp : 0.0f;
            if (su) su[o] = 0.0f;
        }
    }
}
