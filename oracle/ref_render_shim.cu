// ref_render_shim.cu -- TEST INFRASTRUCTURE ONLY.  extern "C" entry points around the UNMODIFIED reference
// renderers (src/density_grid_renderer.cu, src/velocity_grid_renderer.cu), compiled where they lie against
// oracle/sfml_stub (a recording stand-in for the SFML types they use).  Pins SURVEY.md 8(f3): what the
// reference's own kernels produce for a given field, as the app would draw it (src/simulation.cpp:67-73).
#include <cstddef>
#include <cstring>

#include <cuda_runtime.h>

#include "density_grid_renderer.cuh"
#include "velocity_grid_renderer.cuh"

extern "C" int ref_render_device_count() {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

// density_grid_renderer::draw (src/density_grid_renderer.cu:38-56): n x n RGBA8 pixels
extern "C" int ref_render_density(size_t n, const float* density, float mr, float mg, float mb, unsigned target_w,
                                  unsigned target_h, unsigned char* rgba_out) {
    grid<float> g(n, n, 0.f);
    std::memcpy(g.data(), density, sizeof(float) * n * n);
    density_grid_renderer renderer(n, n);
    sf::RenderTarget target;
    target.m_size = sf::Vector2u{target_w, target_h};
    renderer.draw(target, g, color_multipliers{mr, mg, mb});
    if (target.m_pixels.size() != n * n * 4) return 1;
    std::memcpy(rgba_out, target.m_pixels.data(), n * n * 4);
    return cudaGetLastError() == cudaSuccess ? 0 : 2;
}

// velocity_grid_renderer::draw (src/velocity_grid_renderer.cu:55-72): n x n segments as (start.x, start.y, end.x, end.y);
// colours_ok reports whether every vertex colour is the constant white the kernel writes (:34-42)
extern "C" int ref_render_velocity(size_t n, const float* u, const float* v, unsigned target_w, unsigned target_h,
                                   float* lines_out, int* colours_ok) {
    grid<float> gu(n, n, 0.f), gv(n, n, 0.f);
    std::memcpy(gu.data(), u, sizeof(float) * n * n);
    std::memcpy(gv.data(), v, sizeof(float) * n * n);
    velocity_grid_renderer renderer(n, n);
    sf::RenderTarget target;
    target.m_size = sf::Vector2u{target_w, target_h};
    renderer.draw(target, gu, gv);
    if (target.m_vertices.size() != 2 * n * n || target.m_type != sf::Lines) return 1;
    int ok = 1;
    for (size_t c = 0; c < n * n; ++c) {
        const sf::Vertex& a = target.m_vertices[2 * c];
        const sf::Vertex& b = target.m_vertices[2 * c + 1];
        lines_out[4 * c + 0] = a.position.x;
        lines_out[4 * c + 1] = a.position.y;
        lines_out[4 * c + 2] = b.position.x;
        lines_out[4 * c + 3] = b.position.y;
        ok &= (a.color.r == 255 && a.color.g == 255 && a.color.b == 255 && a.color.a == 255 && b.color.r == 255 &&
               b.color.g == 255 && b.color.b == 255 && b.color.a == 255);
    }
    if (colours_ok) *colours_ok = ok;
    return cudaGetLastError() == cudaSuccess ? 0 : 2;
}
