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
};

class simulation_headless {
public:
    explicit simulation_headless(simulation_config const& config)
        : m_config(config),
          m_density_grid{config.height, config.width, 0.f},
          m_horizontal_velocity_grid{config.height, config.width, 0.f},
          m_vertical_velocity_grid{config.height, config.width, 0.f},
          m_horizontal_velocity_source_grid{config.height, config.width, 0.f},
          m_vertical_velocity_source_grid{config.height, config.width, 0.f},
          m_density_source_grid{config.height, config.width, 0.f} {
        switch (config.solver) {
            case solver_type::b200:
                m_solver = std::make_unique<fluid_solver_b200>(config.height, config.width, config.solver_options);
                break;
            case solver_type::b200_cpu_exact: {
                fluid_solver_b200::options o = fluid_solver_b200::options::cpu_compatible();
                o.device = config.solver_options.device;
                m_solver = std::make_unique<fluid_solver_b200>(config.height, config.width, o);
                break;
            }
        }
    }

    void reset() {
        std::fill(m_density_grid.begin(), m_density_grid.end(), 0.f);
        std::fill(m_horizontal_velocity_grid.begin(), m_horizontal_velocity_grid.end(), 0.f);
        std::fill(m_vertical_velocity_grid.begin(), m_vertical_velocity_grid.end(), 0.f);
    }

    // the mouse -> cell mapping of src/simulation.cpp:36-42
    bool coordinates_to_cell(float const x, float const y, unsigned const target_w, unsigned const target_h, size_t& i,
                             size_t& j) const {
        return ::coordinates_to_cell(m_config.height, m_config.width, x, y, target_w, target_h, i, j);
    }

    void add_density_source(size_t const i, size_t const j, float const value) {
        m_density_source_grid(i, j) += value * m_config.width * m_config.height;
    }

    void add_velocity_source(size_t const i, size_t const j, float const horizontal_value, float const vertical_value) {
        m_horizontal_velocity_source_grid(i, j) += horizontal_value * m_config.width * m_config.height;
        m_vertical_velocity_source_grid(i, j) += vertical_value * m_config.width * m_config.height;
    }

    void update(std::chrono::duration<float> const& dt) {
        m_solver->solve(m_density_grid, m_density_source_grid, m_config.diffusion_rate, m_horizontal_velocity_grid,
                        m_vertical_velocity_grid, m_horizontal_velocity_source_grid, m_vertical_velocity_source_grid,
                        m_config.viscosity, dt.count());
        std::fill(m_density_source_grid.begin(), m_density_source_grid.end(), 0.0f);
        std::fill(m_horizontal_velocity_source_grid.begin(), m_horizontal_velocity_source_grid.end(), 0.0f);
        std::fill(m_vertical_velocity_source_grid.begin(), m_vertical_velocity_source_grid.end(), 0.0f);
    }

    // density_grid_renderer::draw without SFML (src/density_grid_renderer.cu:38-56): the density image of the
    // current state as a binary PPM (P6); colour = clamp(multiplier * density, 0, 255) per channel
    void draw_density_ppm(std::string const& path, float r = 255.f, float g = 255.f, float b = 255.f) {
        auto* solver = dynamic_cast<fluid_solver_b200*>(m_solver.get());
        if (!solver) throw std::runtime_error("draw_density_ppm needs the b200 solver");
        std::vector<unsigned char> rgba(m_config.width * m_config.height * 4);
        solver->render_density_rgba(r, g, b, rgba.data());
        FILE* f = std::fopen(path.c_str(), "wb");
        if (!f) throw std::runtime_error("cannot open " + path);
        std::fprintf(f, "P6\n%zu %zu\n255\n", m_config.width, m_config.height);
        for (size_t p = 0; p < m_config.width * m_config.height; ++p) std::fwrite(&rgba[4 * p], 1, 3, f);
        std::fclose(f);
    }

    grid<float>& density() { return m_density_grid; }
    grid<float>& horizontal_velocity() { return m_horizontal_velocity_grid; }
    grid<float>& vertical_velocity() { return m_vertical_velocity_grid; }
    simulation_config const& config() const { return m_config; }

private:
    simulation_config m_config;
    grid<float> m_density_grid;
    grid<float> m_horizontal_velocity_grid;
    grid<float> m_vertical_velocity_grid;
    grid<float> m_horizontal_velocity_source_grid;
    grid<float> m_vertical_velocity_source_grid;
    grid<float> m_density_source_grid;
    std::unique_ptr<fluid_solver> m_solver;
};
