// ref_cpu_shim.cpp -- TEST INFRASTRUCTURE ONLY (see oracle/README in DESIGN.md section "Oracle").
//
// extern "C" entry points around the UNMODIFIED reference CPU solver.  The reference
// translation unit /root/reference/src/fluid_solver_cpu.cpp is compiled where it lies
// (oracle/Makefile); nothing of it is copied here.  The private stage methods
// (src/fluid_solver_cpu.hpp:22-49) are reached with the access-specifier trick of
// SURVEY.md Appendix E so that the solve() sequence (src/fluid_solver_cpu.cpp:15-30)
// can be replayed with arbitrary iteration counts.
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <functional>
#include <vector>

#define private public
#include "fluid_solver_cpu.hpp"
#undef private

namespace {
grid<float> to_grid(const float* p, size_t n) {
    grid<float> g(n, n, 0.f);
    std::memcpy(g.data(), p, sizeof(float) * n * n);
    return g;
}
void from_grid(float* p, grid<float> const& g, size_t n) { std::memcpy(p, g.data(), sizeof(float) * n * n); }

using bnd_fn = void (*)(grid<float>&);
bnd_fn bnd_of(int kind) {
    switch (kind) {
        case 1: return &fluid_solver_cpu::set_boundary_opposite_horizontal;
        case 2: return &fluid_solver_cpu::set_boundary_opposite_vertical;
        default: return &fluid_solver_cpu::set_boundary_continuous;
    }
}
}  // namespace

extern "C" {

// fluid_solver_cpu::solve, literally (K = 20 everywhere), `steps` times.
void ref_cpu_solve(size_t n, float* d, const float* sd, float diffusion_rate, float* u, float* v,
                   const float* su, const float* sv, float viscosity, float dt, unsigned steps) {
    grid<float> gd = to_grid(d, n), gu = to_grid(u, n), gv = to_grid(v, n);
    grid<float> gsd = to_grid(sd, n), gsu = to_grid(su, n), gsv = to_grid(sv, n);
    fluid_solver_cpu s;
    for (unsigned k = 0; k < steps; ++k) s.solve(gd, gsd, diffusion_rate, gu, gv, gsu, gsv, viscosity, dt);
    from_grid(d, gd, n);
    from_grid(u, gu, n);
    from_grid(v, gv, n);
}

// solve()'s stage sequence (src/fluid_solver_cpu.cpp:15-30) with free iteration counts.
void ref_cpu_step_k(size_t n, float* d, const float* sd, float diffusion_rate, float* u, float* v,
                    const float* su, const float* sv, float viscosity, float dt, unsigned kd,
                    unsigned kp, unsigned steps) {
    grid<float> gd = to_grid(d, n), gu = to_grid(u, n), gv = to_grid(v, n);
    grid<float> gsd = to_grid(sd, n), gsu = to_grid(su, n), gsv = to_grid(sv, n);
    fluid_solver_cpu s;
    for (unsigned k = 0; k < steps; ++k) {
        s.add_sources(gd, gsd, dt);
        s.diffuse(gd, &fluid_solver_cpu::set_boundary_continuous, diffusion_rate, dt, kd);
        s.advect(gd, gu, gv, &fluid_solver_cpu::set_boundary_continuous, dt, true);
        s.add_sources(gu, gsu, dt);
        s.add_sources(gv, gsv, dt);
        s.diffuse(gu, &fluid_solver_cpu::set_boundary_opposite_horizontal, viscosity, dt, kd);
        s.diffuse(gv, &fluid_solver_cpu::set_boundary_opposite_vertical, viscosity, dt, kd);
        s.project(gu, gv, kp);
        grid<float> tu = gu, tv = gv;
        s.advect(gu, tu, tv, &fluid_solver_cpu::set_boundary_opposite_horizontal, dt, false);
        s.advect(gv, tu, tv, &fluid_solver_cpu::set_boundary_opposite_vertical, dt, false);
        s.project(gu, gv, kp);
    }
    from_grid(d, gd, n);
    from_grid(u, gu, n);
    from_grid(v, gv, n);
}

void ref_cpu_set_bnd(size_t n, float* f, int kind) {
    grid<float> g = to_grid(f, n);
    bnd_of(kind)(g);
    from_grid(f, g, n);
}

void ref_cpu_add_sources(size_t n, float* f, const float* s, float dt) {
    grid<float> g = to_grid(f, n), gs = to_grid(s, n);
    fluid_solver_cpu sol;
    sol.add_sources(g, gs, dt);
    from_grid(f, g, n);
}

void ref_cpu_diffuse(size_t n, float* f, int kind, float rate, float dt, unsigned iters) {
    grid<float> g = to_grid(f, n);
    fluid_solver_cpu sol;
    sol.diffuse(g, bnd_of(kind), rate, dt, iters);
    from_grid(f, g, n);
}

void ref_cpu_advect(size_t n, float* f, const float* u, const float* v, int kind, float dt, int trace) {
    grid<float> g = to_grid(f, n), gu = to_grid(u, n), gv = to_grid(v, n);
    fluid_solver_cpu sol;
    sol.advect(g, gu, gv, bnd_of(kind), dt, trace != 0);
    from_grid(f, g, n);
}

void ref_cpu_project(size_t n, float* u, float* v, unsigned iters) {
    grid<float> gu = to_grid(u, n), gv = to_grid(v, n);
    fluid_solver_cpu sol;
    sol.project(gu, gv, iters);
    from_grid(u, gu, n);
    from_grid(v, gv, n);
}

}  // extern "C"
