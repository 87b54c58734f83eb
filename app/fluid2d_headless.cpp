// fluid2d_headless -- the reference's app loop without a window.
// What src/app.cpp does per 20 ms frame (fixed timestep :8, mouse -> sources :100-118, simulation.update
// :120) driven by a scripted "mouse": a density emitter and a velocity stirrer moving on circles.
// The reference's parse_simulation_config ignores argv ("TODO", src/app.cpp:27-36); this one parses it.
//
//   fluid2d_headless [--size N | --width W --height H] [--steps S] [--dt 0.02] [--diffusion 0.5]
//                    [--viscosity 1e-6] [--kd 15] [--kp 20] [--no-smooth] [--exact-divide] [--solver b200|b200-cpu-exact]
//                    [--load prefix] [--dump prefix] [--dump-every K] [--ppm file.ppm] [--quiet]
//   --load/--dump use prefix_{density,u,v}.npy (include/f2d_npy.hpp)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "f2d_npy.hpp"
#include "simulation_headless.hpp"

namespace {
struct args {
    size_t width = 256, height = 256;
    int steps = 100, dump_every = 0;
    float dt = 0.02f;
    bool quiet = false;
    std::string load, dump, ppm;
    simulation_config cfg;
};

bool parse(int argc, char** argv, args& a) {
    for (int i = 1; i < argc; ++i) {
        auto is = [&](const char* k) { return std::strcmp(argv[i], k) == 0; };
        auto val = [&]() -> const char* { return (i + 1 < argc) ? argv[++i] : nullptr; };
        const char* v = nullptr;
        if (is("--size")) { if (!(v = val())) return false; a.width = a.height = std::strtoul(v, nullptr, 10); }
        else if (is("--width")) { if (!(v = val())) return false; a.width = std::strtoul(v, nullptr, 10); }
        else if (is("--height")) { if (!(v = val())) return false; a.height = std::strtoul(v, nullptr, 10); }
        else if (is("--steps")) { if (!(v = val())) return false; a.steps = std::atoi(v); }
        else if (is("--dt")) { if (!(v = val())) return false; a.dt = std::strtof(v, nullptr); }
        else if (is("--diffusion")) { if (!(v = val())) return false; a.cfg.diffusion_rate = std::strtof(v, nullptr); }
        else if (is("--viscosity")) { if (!(v = val())) return false; a.cfg.viscosity = std::strtof(v, nullptr); }
        else if (is("--kd")) { if (!(v = val())) return false; a.cfg.solver_options.diffuse_iterations = std::atoi(v); }
        else if (is("--kp")) { if (!(v = val())) return false; a.cfg.solver_options.project_iterations = std::atoi(v); }
        else if (is("--no-smooth")) a.cfg.solver_options.smooth = false;
        else if (is("--exact-divide")) a.cfg.solver_options.exact_divide = true;
        else if (is("--solver")) {  // the switch of src/simulation.cpp:17-26
            if (!(v = val())) return false;
            if (std::strcmp(v, "b200") == 0) a.cfg.solver = solver_type::b200;
            else if (std::strcmp(v, "b200-cpu-exact") == 0) a.cfg.solver = solver_type::b200_cpu_exact;
            else return false;
        }
        else if (is("--load")) { if (!(v = val())) return false; a.load = v; }
        else if (is("--dump")) { if (!(v = val())) return false; a.dump = v; }
        else if (is("--ppm")) { if (!(v = val())) return false; a.ppm = v; }
        else if (is("--dump-every")) { if (!(v = val())) return false; a.dump_every = std::atoi(v); }
        else if (is("--quiet")) a.quiet = true;
        else return false;
    }
    a.cfg.width = a.width;
    a.cfg.height = a.height;
    return a.width >= 3 && a.height >= 3 && a.steps >= 0;
}

void load_field(const std::string& path, grid<float>& g) {
    std::vector<float> data;
    size_t r = 0, c = 0;
    f2d_npy::load(path, data, r, c);
    if (r != g.rows() || c != g.cols()) throw std::runtime_error(path + ": shape does not match the grid");
    std::copy(data.begin(), data.end(), g.begin());
}

void dump(simulation_headless& sim, const std::string& prefix) {
    f2d_npy::save(prefix + "_density.npy", sim.density().data(), sim.density().rows(), sim.density().cols());
    f2d_npy::save(prefix + "_u.npy", sim.horizontal_velocity().data(), sim.density().rows(), sim.density().cols());
    f2d_npy::save(prefix + "_v.npy", sim.vertical_velocity().data(), sim.density().rows(), sim.density().cols());
}
}  // namespace

int main(int argc, char** argv) {
    args a;
    if (!parse(argc, argv, a)) {
        std::fprintf(stderr, "usage: %s [--size N] [--steps S] [--dt s] [--diffusion r] [--viscosity v] [--kd K] [--kp K]\n"
                             "          [--no-smooth] [--exact-divide] [--solver b200|b200-cpu-exact] [--load prefix] [--dump prefix] [--dump-every K] [--quiet]\n", argv[0]);
        return 64;
    }
    try {
        simulation_headless sim(a.cfg);
        if (!a.load.empty()) {
            load_field(a.load + "_density.npy", sim.density());
            load_field(a.load + "_u.npy", sim.horizontal_velocity());
            load_field(a.load + "_v.npy", sim.vertical_velocity());
        }
        const double two_pi = 6.283185307179586;
        const auto t0 = std::chrono::steady_clock::now();
        for (int s = 0; s < a.steps; ++s) {
            // scripted mouse: left button = density (src/app.cpp:103-107), right button drag = velocity (:110-118)
            const double ph = two_pi * s / 97.0;
            const size_t ci = static_cast<size_t>(a.height * (0.5 + 0.25 * std::sin(ph)));
            const size_t cj = static_cast<size_t>(a.width * (0.5 + 0.25 * std::cos(ph)));
            sim.add_density_source(ci, cj, 0.075f);
            const size_t vi = static_cast<size_t>(a.height * (0.5 + 0.3 * std::sin(-1.7 * ph)));
            const size_t vj = static_cast<size_t>(a.width * (0.5 + 0.3 * std::cos(-1.7 * ph)));
            sim.add_velocity_source(vi, vj, 0.05f * static_cast<float>(-std::sin(-1.7 * ph)), 0.05f * static_cast<float>(std::cos(-1.7 * ph)));
            sim.update(std::chrono::duration<float>(a.dt));
            if (!a.dump.empty() && a.dump_every > 0 && (s + 1) % a.dump_every == 0) dump(sim, a.dump + "_step" + std::to_string(s + 1));
        }
        const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (!a.dump.empty()) dump(sim, a.dump);
        if (!a.ppm.empty()) sim.draw_density_ppm(a.ppm, 255.f, 160.f, 64.f);
        double sum = 0, umax = 0;
        for (auto it = sim.density().cbegin(); it != sim.density().cend(); ++it) sum += *it;
        for (auto it = sim.horizontal_velocity().cbegin(); it != sim.horizontal_velocity().cend(); ++it) umax = std::max(umax, std::fabs(static_cast<double>(*it)));
        if (!a.quiet)
            std::printf("fluid2d_headless: %zux%zu, %d steps, %.3f ms/step through fluid_solver::solve, density sum %.9e, max|u| %.9e\n",
                        a.width, a.height, a.steps, a.steps ? 1e3 * sec / a.steps : 0.0, sum, umax);
    } catch (std::exception const& e) {
        std::fprintf(stderr, "fluid2d_headless: %s\n", e.what());
        return 2;
    }
    return 0;
}
