// headless_sim.cpp -- the reference's caller contract without SFML: what simulation::update does
// around fluid_solver::solve (src/simulation.cpp:17-65), driving fluid_solver_b200 through the
// reference interface.  Sources are injected scaled by width*height (src/simulation.cpp:44-51) and
// zeroed after every solve (src/simulation.cpp:62-64).
//
//   g++ -std=c++14 -O2 -Iinclude examples/headless_sim.cpp -Lfluid-2d_b200 -lf2d -Wl,-rpath,$PWD/fluid-2d_b200 -o headless_sim
//   ./headless_sim [N=256] [steps=100]
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <memory>

#include "fluid_solver_b200.hpp"

int main(int argc, char** argv) {
    const size_t n = argc > 1 ? static_cast<size_t>(atol(argv[1])) : 256;
    const int steps = argc > 2 ? atoi(argv[2]) : 100;
    const float dt = 0.02f, diffusion_rate = 0.5f, viscosity = 1e-6f;  // src/app.cpp:8,33-34

    grid<float> density(n, n, 0.f), u(n, n, 0.f), v(n, n, 0.f);
    grid<float> density_source(n, n, 0.f), u_source(n, n, 0.f), v_source(n, n, 0.f);
    std::unique_ptr<fluid_solver> solver;
    try {
        solver.reset(new fluid_solver_b200(n, n));  // the new case of the switch at src/simulation.cpp:17-26
    } catch (std::exception const& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 2;
    }
    const auto t0 = std::chrono::steady_clock::now();
    for (int s = 0; s < steps; ++s) {
        // what app::update does for a pressed mouse button (src/app.cpp:106,115-117)
        const size_t ci = n / 4, cj = n / 2;
        density_source(ci, cj) += 0.075f * n * n;
        v_source(ci, cj) += 0.05f * 1.0f * n * n;
        solver->solve(density, density_source, diffusion_rate, u, v, u_source, v_source, viscosity, dt);
        std::fill(density_source.begin(), density_source.end(), 0.f);
        std::fill(u_source.begin(), u_source.end(), 0.f);
        std::fill(v_source.begin(), v_source.end(), 0.f);
    }
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    double sum = 0, vmax = 0;
    for (auto it = density.cbegin(); it != density.cend(); ++it) sum += *it;
    for (auto it = v.cbegin(); it != v.cend(); ++it) vmax = std::max(vmax, static_cast<double>(*it));
    std::printf("headless_sim: %zux%zu, %d steps through fluid_solver::solve, %.3f ms/step, density sum %.6e, max v %.6e\n",
                n, n, steps, 1e3 * sec / steps, sum, vmax);
    return (sum > 0 && vmax > 0) ? 0 : 1;
}
