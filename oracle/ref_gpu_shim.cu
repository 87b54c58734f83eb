// ref_gpu_shim.cu -- TEST INFRASTRUCTURE ONLY.
//
// extern "C" entry points around the UNMODIFIED reference GPU solver
// (/root/reference/src/fluid_solver_gpu.cu, compiled where it lies by oracle/Makefile with
// nvcc defaults for sm_100a -- the defaults define the oracle's arithmetic).  Used on the GPU
// box (a) to produce the golden fixtures tests/golden/refgpu_*.npz that pin the SFO_SEM_GPU
// restatement, and (b) live in `-m gpu` tests / bench.py as "the kernel to beat".
// Private stage methods and buffers (src/fluid_solver_gpu.cuh:27-69) are reached with the
// access-specifier trick of SURVEY.md Appendix E.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <functional>
#include <vector>

#define private public
#include "fluid_solver_gpu.cuh"
#undef private
#include "utilities.hpp"

namespace {
grid<float> to_grid(const float* p, size_t n) {
    grid<float> g(n, n, 0.f);
    std::memcpy(g.data(), p, sizeof(float) * n * n);
    return g;
}
void from_grid(float* p, grid<float> const& g, size_t n) { std::memcpy(p, g.data(), sizeof(float) * n * n); }

using bnd_fn = void (*)(linear_buffer<float>&, size_t, size_t);
bnd_fn bnd_of(int kind) {
    switch (kind) {
        case 1: return &fluid_solver_gpu::set_boundary_opposite_horizontal;
        case 2: return &fluid_solver_gpu::set_boundary_opposite_vertical;
        default: return &fluid_solver_gpu::set_boundary_continuous;
    }
}
}  // namespace

extern "C" {

int ref_gpu_device_count() {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

// fluid_solver_gpu::solve, literally (Kd=15, Kp=20, smooth on), `steps` times.
// Returns elapsed host milliseconds of the solve loop (excluding grid set-up).
double ref_gpu_solve(size_t n, float* d, const float* sd, float diffusion_rate, float* u, float* v,
                     const float* su, const float* sv, float viscosity, float dt, unsigned steps) {
    grid<float> gd = to_grid(d, n), gu = to_grid(u, n), gv = to_grid(v, n);
    grid<float> gsd = to_grid(sd, n), gsu = to_grid(su, n), gsv = to_grid(sv, n);
    fluid_solver_gpu s(n, n);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (unsigned k = 0; k < steps; ++k) s.solve(gd, gsd, diffusion_rate, gu, gv, gsu, gsv, viscosity, dt);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    from_grid(d, gd, n);
    from_grid(u, gu, n);
    from_grid(v, gv, n);
    return (double)ms;
}

// solve()'s stage sequence (src/fluid_solver_gpu.cu:232-257) with free iteration counts and an
// optional smooth; host<->device copies as in solve().  Returns elapsed ms.
double ref_gpu_step_k(size_t n, float* d, const float* sd, float diffusion_rate, float* u, float* v,
                      const float* su, const float* sv, float viscosity, float dt, unsigned kd,
                      unsigned kp, int do_smooth, unsigned steps) {
    grid<float> gd = to_grid(d, n), gu = to_grid(u, n), gv = to_grid(v, n);
    grid<float> gsd = to_grid(sd, n), gsu = to_grid(su, n), gsv = to_grid(sv, n);
    fluid_solver_gpu s(n, n);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (unsigned k = 0; k < steps; ++k) {
        copy(s.m_density_buffer, gd);
        copy(s.m_horizontal_velocity_buffer, gu);
        copy(s.m_vertical_velocity_buffer, gv);
        s.add_sources(s.m_density_buffer, gsd, dt);
        s.diffuse(s.m_density_buffer, &fluid_solver_gpu::set_boundary_continuous, diffusion_rate, dt, kd);
        s.advect(s.m_density_buffer, s.m_horizontal_velocity_buffer, s.m_vertical_velocity_buffer,
                 &fluid_solver_gpu::set_boundary_continuous, dt, true);
        if (do_smooth) s.smooth(s.m_density_buffer);
        s.add_sources(s.m_horizontal_velocity_buffer, gsu, dt);
        s.add_sources(s.m_vertical_velocity_buffer, gsv, dt);
        s.diffuse(s.m_horizontal_velocity_buffer, &fluid_solver_gpu::set_boundary_opposite_horizontal, viscosity, dt, kd);
        s.diffuse(s.m_vertical_velocity_buffer, &fluid_solver_gpu::set_boundary_opposite_vertical, viscosity, dt, kd);
        s.project(s.m_horizontal_velocity_buffer, s.m_vertical_velocity_buffer, kp);
        copy(s.m_temp_buffer_2, s.m_horizontal_velocity_buffer);
        copy(s.m_temp_buffer_3, s.m_vertical_velocity_buffer);
        s.advect(s.m_horizontal_velocity_buffer, s.m_temp_buffer_2, s.m_temp_buffer_3,
                 &fluid_solver_gpu::set_boundary_opposite_horizontal, dt, false);
        s.advect(s.m_vertical_velocity_buffer, s.m_temp_buffer_2, s.m_temp_buffer_3,
                 &fluid_solver_gpu::set_boundary_opposite_vertical, dt, false);
        s.project(s.m_horizontal_velocity_buffer, s.m_vertical_velocity_buffer, kp);
        copy(gd, s.m_density_buffer);
        copy(gu, s.m_horizontal_velocity_buffer);
        copy(gv, s.m_vertical_velocity_buffer);
    }
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    from_grid(d, gd, n);
    from_grid(u, gu, n);
    from_grid(v, gv, n);
    return (double)ms;
}

// ---- single stages on host arrays (upload, run the reference stage, download) ----
void ref_gpu_set_bnd(size_t n, float* f, int kind) {
    fluid_solver_gpu s(n, n);
    grid<float> g = to_grid(f, n);
    copy(s.m_density_buffer, g);
    bnd_of(kind)(s.m_density_buffer, n, n);
    copy(g, s.m_density_buffer);
    from_grid(f, g, n);
}

void ref_gpu_add_sources(size_t n, float* f, const float* src, float dt) {
    fluid_solver_gpu s(n, n);
    grid<float> g = to_grid(f, n), gs = to_grid(src, n);
    copy(s.m_density_buffer, g);
    s.add_sources(s.m_density_buffer, gs, dt);
    copy(g, s.m_density_buffer);
    from_grid(f, g, n);
}

void ref_gpu_diffuse(size_t n, float* f, int kind, float rate, float dt, unsigned iters) {
    fluid_solver_gpu s(n, n);
    grid<float> g = to_grid(f, n);
    copy(s.m_density_buffer, g);
    s.diffuse(s.m_density_buffer, bnd_of(kind), rate, dt, iters);
    copy(g, s.m_density_buffer);
    from_grid(f, g, n);
}

void ref_gpu_smooth(size_t n, float* f) {
    fluid_solver_gpu s(n, n);
    grid<float> g = to_grid(f, n);
    copy(s.m_density_buffer, g);
    s.smooth(s.m_density_buffer);
    copy(g, s.m_density_buffer);
    from_grid(f, g, n);
}

// advect f with velocities (u,v); trace!=0 is the density scatter.  (u,v) are uploaded into the
// velocity buffers, which never alias the advected buffer or temp_buffer_1.
void ref_gpu_advect(size_t n, float* f, const float* u, const float* v, int kind, float dt, int trace) {
    fluid_solver_gpu s(n, n);
    grid<float> g = to_grid(f, n), gu = to_grid(u, n), gv = to_grid(v, n);
    copy(s.m_density_buffer, g);
    copy(s.m_horizontal_velocity_buffer, gu);
    copy(s.m_vertical_velocity_buffer, gv);
    s.advect(s.m_density_buffer, s.m_horizontal_velocity_buffer, s.m_vertical_velocity_buffer, bnd_of(kind), dt, trace != 0);
    copy(g, s.m_density_buffer);
    from_grid(f, g, n);
}

// project; optionally returns the final pressure (temp_buffer_2) and divergence (temp_buffer_1).
void ref_gpu_project(size_t n, float* u, float* v, unsigned iters, float* p_out, float* div_out) {
    fluid_solver_gpu s(n, n);
    grid<float> gu = to_grid(u, n), gv = to_grid(v, n);
    copy(s.m_horizontal_velocity_buffer, gu);
    copy(s.m_vertical_velocity_buffer, gv);
    s.project(s.m_horizontal_velocity_buffer, s.m_vertical_velocity_buffer, iters);
    copy(gu, s.m_horizontal_velocity_buffer);
    copy(gv, s.m_vertical_velocity_buffer);
    from_grid(u, gu, n);
    from_grid(v, gv, n);
    if (p_out) cudaMemcpy(p_out, s.m_temp_buffer_2.data(), sizeof(float) * n * n, cudaMemcpyDeviceToHost);
    if (div_out) cudaMemcpy(div_out, s.m_temp_buffer_1.data(), sizeof(float) * n * n, cudaMemcpyDeviceToHost);
}

}  // extern "C"
