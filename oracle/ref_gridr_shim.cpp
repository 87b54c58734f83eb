// ref_gridr_shim.cpp -- TEST INFRASTRUCTURE ONLY.  grid_renderer::coordinates_to_cell (src/grid_renderer.cpp:3-14),
// the mouse -> cell mapping of the reference app (src/app.cpp:100-118), compiled UNMODIFIED against oracle/sfml_stub.
#include <cstddef>

#include "grid_renderer.hpp"

namespace {
struct probe : grid_renderer {  // the constructor is protected
    probe(size_t rows, size_t cols) : grid_renderer(rows, cols) {}
};
}  // namespace

extern "C" int ref_coordinates_to_cell(size_t rows, size_t cols, float x, float y, unsigned target_w, unsigned target_h,
                                       size_t* i, size_t* j) {
    probe p(rows, cols);
    sf::RenderTarget target;
    target.m_size = sf::Vector2u{target_w, target_h};
    return p.coordinates_to_cell(x, y, target, *i, *j) ? 1 : 0;
}
