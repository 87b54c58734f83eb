// fluid_solver_b200.hpp -- header-only C++ adapter: the B200 solver behind the reference's
// `fluid_solver` interface, so that it drops in next to fluid_solver_cpu / fluid_solver_gpu
// (factory switch at src/simulation.cpp:17-26).  All work happens in libf2d.so through the C ABI
// (include/f2d.h); this class only forwards.
//
//   fluid_solver_b200(rows, cols)        mirrors fluid_solver_gpu(rows, cols)
//                                        (src/fluid_solver_gpu.cu:209-218, called with
//                                        (config.height, config.width) at src/simulation.cpp:22)
//   solve(...)                           mirrors fluid_solver_gpu::solve (src/fluid_solver_gpu.cu:222-258)
//
// Error behaviour: the reference's solve() returns void and discards every CUDA status.  This
// adapter keeps the void signature; a failing call throws std::runtime_error carrying
// f2d_last_error() (construction without a GPU fails loudly -- there is no CPU fallback).
// Beyond the interface, step()/upload()/download() expose the device-resident extension.
#pragma once

#include <stdexcept>
#include <string>

#include "f2d.h"
#include "fluid_solver.hpp"

class fluid_solver_b200 : public fluid_solver {
public:
    struct options {
        unsigned diffuse_iterations = 15;  // src/fluid_solver_gpu.cu:238,245-246
        unsigned project_iterations = 20;  // src/fluid_solver_gpu.cu:247,252
        bool smooth = true;                // src/fluid_solver_gpu.cu:240
        bool exact_divide = false;         // true: the reference's fp64 divide in diffuse (bit-exact u, v)
        bool cpu_semantics = false;        // true: fluid_solver_cpu's arithmetic (F2D_SEM_CPU), bit-identical to it
        int device = -1;

        // the configuration that reproduces fluid_solver_cpu::solve bit for bit (src/fluid_solver_cpu.cpp:15-30)
        static options cpu_compatible() {
            options o;
            o.diffuse_iterations = o.project_iterations = 20;
            o.smooth = false;
            o.cpu_semantics = true;
            return o;
        }
    };

    fluid_solver_b200(size_t const rows, size_t const cols) : fluid_solver_b200(rows, cols, options{}) {}

    fluid_solver_b200(size_t const rows, size_t const cols, options const& opt) : rows_(rows), cols_(cols) {
        f2d_config cfg;
        check(f2d_config_default(&cfg, static_cast<uint32_t>(rows), static_cast<uint32_t>(cols)));
        cfg.diffuse_iters = opt.diffuse_iterations;
        cfg.project_iters = opt.project_iterations;
        cfg.smooth = opt.smooth ? 1u : 0u;
        cfg.divide_mode = opt.exact_divide ? F2D_DIV_F64 : F2D_DIV_F32_CORR;
        cfg.semantics = opt.cpu_semantics ? F2D_SEM_CPU : F2D_SEM_GPU;
        cfg.device = opt.device;
        check(f2d_create(&cfg, &handle_));
    }

    ~fluid_solver_b200() override { f2d_destroy(handle_); }

    fluid_solver_b200(fluid_solver_b200 const&) = delete;
    fluid_solver_b200& operator=(fluid_solver_b200 const&) = delete;

    void solve(grid<float>& density_grid,
               grid<float> const& density_source_grid,
               float const diffusion_rate,
               grid<float>& horizontal_velocity_grid,
               grid<float>& vertical_velocity_grid,
               grid<float> const& horizontal_velocity_source_grid,
               grid<float> const& vertical_velocity_source_grid,
               float const viscosity,
               float const dt) override {
        check(f2d_solve_host(handle_, density_grid.data(), density_source_grid.data(), diffusion_rate,
                             horizontal_velocity_grid.data(), vertical_velocity_grid.data(),
                             horizontal_velocity_source_grid.data(), vertical_velocity_source_grid.data(),
                             viscosity, dt));
    }

    // ---- device-resident extension (no per-step PCIe traffic)
    void upload(grid<float> const& d, grid<float> const& u, grid<float> const& v) {
        check(f2d_upload(handle_, d.data(), u.data(), v.data()));
    }
    void set_sources(grid<float> const& sd, grid<float> const& su, grid<float> const& sv) {
        check(f2d_set_sources(handle_, sd.data(), su.data(), sv.data()));
    }
    void clear_sources() { check(f2d_clear_sources(handle_)); }
    void step(float diffusion_rate, float viscosity, float dt, unsigned nsteps = 1) {
        check(f2d_step(handle_, diffusion_rate, viscosity, dt, nsteps));
    }
    float step_timed(float diffusion_rate, float viscosity, float dt, unsigned nsteps) {
        float ms = 0.f;
        check(f2d_step_timed(handle_, diffusion_rate, viscosity, dt, nsteps, &ms));
        return ms;
    }
    void download(grid<float>& d, grid<float>& u, grid<float>& v) {
        check(f2d_download(handle_, d.data(), u.data(), v.data()));
    }
    // headless renderers on the device-resident state of the last solve()/step()
    // (grid_to_image_kernel, src/density_grid_renderer.cu:10-29; velocity_to_lines_kernel, src/velocity_grid_renderer.cu:8-44)
    void render_density_rgba(float r, float g, float b, unsigned char* rgba) {
        check(f2d_render_density_rgba(handle_, r, g, b, rgba));
    }
    void render_velocity_lines(float horizontal_scale, float vertical_scale, float* lines) {
        check(f2d_render_velocity_lines(handle_, horizontal_scale, vertical_scale, lines));
    }
    // Page-lock a caller-owned grid so that solve() moves it by DMA at PCIe speed instead of through the driver's
    // staging copy (f2d_pin_host).  The grid must stay alive, and must not be resized, until unpin() or until this
    // solver is destroyed -- which is why solve() never does this behind the caller's back.
    void pin(grid<float> const& g) { check(f2d_pin_host(handle_, g.data(), g.rows() * g.cols() * sizeof(float))); }
    void unpin(grid<float> const& g) { check(f2d_unpin_host(handle_, g.data())); }
    f2d_solver* handle() { return handle_; }

private:
    static void check(int rc) {
        if (rc != F2D_OK) throw std::runtime_error(std::string("fluid_solver_b200: ") + f2d_last_error());
    }
    size_t rows_, cols_;
    f2d_solver* handle_ = nullptr;
};
