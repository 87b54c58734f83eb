// simulation_headless.hpp -- the reference's `simulation` class without SFML.
// Mirrors the caller contract around fluid_solver::solve: six host grids owned by the simulation
// (src/simulation.hpp:64-77), source injection scaled by width*height (src/simulation.cpp:44-51),
// update() = solve() then zero the sources (src/simulation.cpp:53-65), reset() zeroes the state
// (src/simulation.cpp:29-34), solver chosen by an enum (src/simulation.hpp:14, src/simulation.cpp:17-26).
// coordinates_to_cell() is grid_renderer's mouse -> cell mapping (src/grid_renderer.cpp:3-14, reached through
// src/simulation.cpp:36-42); draw() is replaced by draw_density_ppm().
#pragma once

#include <algorithm>
#include <chrono>
#include <memory>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "fluid_solver.hpp"
#include "fluid_solver_b200.hpp"

// The reference's enumerators are { gpu, cpu } (src/simulation.hpp:14); see INTEGRATION.md for the merged switch.
//   b200            the arithmetic of fluid_solver_gpu (the reference's default solver, src/app.cpp:32)
//   b200_cpu_exact  the arithmetic of fluid_solver_cpu, bit for bit (F2D_SEM_CPU: Gauss-Seidel, 20 iterations, no smooth)
enum class solver_type { b200, b200_cpu_exact };

// grid_renderer::coordinates_to_cell (src/grid_renderer.cpp:3-14) without SFML: target position (x, y) on a target of
// target_w x target_h pixels -> grid cell (i, j); false outside the target.  The grid is taken to be stretched over
// the target: float quotient, float product with the size_t extent, truncation -- as the reference computes it.
inline bool coordinates_to_cell(size_t const rows, size_t const cols, float const x, float const y, unsigned const target_w,
                                unsigned const target_h, size_t& i, size_t& j) {
    if (x < 0.f || y < 0.f || x >= target_w || y >= target_h) return false;
    i = static_cast<size_t>(rows * (y / target_h));
    j = static_cast<size_t>(cols * (x / target_w));
    return true;
}

struct simulation_config {
    size_t width = 800;             // src/app.cpp:30-31
    size_t height = 800;
    solver_type solver = solver_type::b200;
    float diffusion_rate = 0.5f;    // src/app.cpp:33
    float viscosity = 1e-6f;        // src/app.cpp:34
    fluid_solver_b200::options solver_options{};
    bool pin_grids = true;          // page-lock the simulation's own grids (f2d_pin_host): they live as long as the solver
};

// State layout: one array of six host grids indexed by the C ABI's field ids (F2D_FIELD_*, include/f2d.h), so the
// same index names a field on the host, in the solver and on the wire; the solver behind it is built by a factory
// keyed on solver_type.  Only the public method names follow the reference's class (they are the caller contract).
class simulation_headless {
public:
    explicit simulation_headless(simulation_config const& config)
        : cfg_(config), solver_(make_solver(config)) {
        fields_.reserve(kFields);
        for (int f = 0; f < kFields; ++f) fields_.emplace_back(config.height, config.width, 0.f);
        if (config.pin_grids)  // page-lock the six grids for the life of this object (they outlive every solve())
            for (auto& g : fields_) solver_->pin(g);
    }

    ~simulation_headless() {
        if (cfg_.pin_grids)
            for (auto& g : fields_) solver_->unpin(g);
    }

    simulation_headless(simulation_headless const&) = delete;
    simulation_headless& operator=(simulation_headless const&) = delete;

    // src/simulation.cpp:29-34: density and velocity go back to zero, pending sources stay
    void reset() {
        for (int f : {F2D_FIELD_DENSITY, F2D_FIELD_U, F2D_FIELD_V}) zero(fields_[f]);
    }

    // the mouse -> cell mapping of src/simulation.cpp:36-42
    bool coordinates_to_cell(float const x, float const y, unsigned const target_w, unsigned const target_h, size_t& i,
                             size_t& j) const {
        return ::coordinates_to_cell(cfg_.height, cfg_.width, x, y, target_w, target_h, i, j);
    }

    // src/simulation.cpp:44-51: injected amounts are scaled by the grid area
    void add_density_source(size_t const i, size_t const j, float const value) { inject(F2D_FIELD_DENSITY_SOURCE, i, j, value); }

    void add_velocity_source(size_t const i, size_t const j, float const horizontal_value, float const vertical_value) {
        inject(F2D_FIELD_U_SOURCE, i, j, horizontal_value);
        inject(F2D_FIELD_V_SOURCE, i, j, vertical_value);
    }

    // src/simulation.cpp:53-65: one solve(), after which the sources count as consumed
    void update(std::chrono::duration<float> const& dt) {
        solver_->solve(fields_[F2D_FIELD_DENSITY], fields_[F2D_FIELD_DENSITY_SOURCE], cfg_.diffusion_rate, fields_[F2D_FIELD_U],
                       fields_[F2D_FIELD_V], fields_[F2D_FIELD_U_SOURCE], fields_[F2D_FIELD_V_SOURCE], cfg_.viscosity, dt.count());
        for (int f : {F2D_FIELD_DENSITY_SOURCE, F2D_FIELD_U_SOURCE, F2D_FIELD_V_SOURCE}) zero(fields_[f]);
    }

    // density_grid_renderer::draw without SFML (src/density_grid_renderer.cu:38-56): the density image of the
    // current state as a binary PPM (P6); colour = clamp(multiplier * density, 0, 255) per channel
    void draw_density_ppm(std::string const& path, float r = 255.f, float g = 255.f, float b = 255.f) {
        size_t const pixels = cfg_.width * cfg_.height;
        std::vector<unsigned char> rgba(pixels * 4), rgb(pixels * 3);
        solver_->render_density_rgba(r, g, b, rgba.data());
        for (size_t p = 0; p < pixels; ++p) std::copy_n(&rgba[4 * p], 3, &rgb[3 * p]);
        FILE* out = std::fopen(path.c_str(), "wb");
        if (!out) throw std::runtime_error("cannot open " + path);
        std::fprintf(out, "P6\n%zu %zu\n255\n", cfg_.width, cfg_.height);
        std::fwrite(rgb.data(), 1, rgb.size(), out);
        std::fclose(out);
    }

    grid<float>& density() { return fields_[F2D_FIELD_DENSITY]; }
    grid<float>& horizontal_velocity() { return fields_[F2D_FIELD_U]; }
    grid<float>& vertical_velocity() { return fields_[F2D_FIELD_V]; }
    grid<float>& field(int f2d_field) { return fields_.at(static_cast<size_t>(f2d_field)); }
    simulation_config const& config() const { return cfg_; }

private:
    static constexpr int kFields = 6;  // F2D_FIELD_DENSITY .. F2D_FIELD_V_SOURCE

    static std::unique_ptr<fluid_solver_b200> make_solver(simulation_config const& c) {
        fluid_solver_b200::options o = c.solver_options;
        if (c.solver == solver_type::b200_cpu_exact) {
            o = fluid_solver_b200::options::cpu_compatible();
            o.device = c.solver_options.device;
        }
        return std::make_unique<fluid_solver_b200>(c.height, c.width, o);
    }

    static void zero(grid<float>& g) { std::fill(g.begin(), g.end(), 0.f); }

    // value * width * height evaluated left to right in float, as the reference's expression rounds
    // (src/simulation.cpp:45, :49-50); a cached float area would round differently
    void inject(int const f, size_t const i, size_t const j, float const value) {
        float const by_width = value * static_cast<float>(cfg_.width);
        fields_[f](i, j) += by_width * static_cast<float>(cfg_.height);
    }

    simulation_config cfg_;
    std::unique_ptr<fluid_solver_b200> solver_;
    std::vector<grid<float>> fields_;
};
