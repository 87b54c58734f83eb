// TEST INFRASTRUCTURE ONLY -- a stand-in for the handful of SFML 2 types the reference's two renderers touch
// (src/density_grid_renderer.{cuh,cu}, src/velocity_grid_renderer.{cuh,cu}, src/grid_renderer.{hpp,cpp}), so that
// those translation units can be compiled UNMODIFIED without SFML or a display and their output captured:
// sf::RenderTarget here records what is drawn instead of rasterising it.  Written for this repo; not SFML code.
// Layouts that the reference's kernels write through (sf::Vertex = position, colour, texture coordinates; 20 bytes)
// follow SFML 2's public headers.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

namespace sf {

typedef std::uint8_t Uint8;

template <typename T>
struct Vector2 {
    T x, y;
};
typedef Vector2<float> Vector2f;
typedef Vector2<unsigned int> Vector2u;

struct Color {
    Uint8 r, g, b, a;
};

struct Vertex {
    Vector2f position;
    Color color;
    Vector2f texCoords;
};

enum PrimitiveType { Points, Lines, LineStrip, Triangles, TriangleStrip, TriangleFan, Quads };

struct FloatRect {
    float left, top, width, height;
};

class Texture {
public:
    bool create(unsigned int width, unsigned int height) {
        m_width = width;
        m_height = height;
        m_pixels.assign((std::size_t)width * height * 4, 0);
        return true;
    }
    void update(const Uint8* pixels, unsigned int width, unsigned int height, unsigned int x, unsigned int y) {
        for (unsigned int r = 0; r < height; ++r)
            for (unsigned int c = 0; c < width * 4; ++c)
                m_pixels[((std::size_t)(y + r) * m_width + x) * 4 + c] = pixels[(std::size_t)r * width * 4 + c];
    }
    unsigned int m_width = 0, m_height = 0;
    std::vector<Uint8> m_pixels;
};

class Sprite {
public:
    explicit Sprite(const Texture& texture) : m_texture(&texture) {}
    void setScale(float x, float y) {
        m_scale_x = x;
        m_scale_y = y;
    }
    FloatRect getLocalBounds() const { return FloatRect{0.f, 0.f, (float)m_texture->m_width, (float)m_texture->m_height}; }
    const Texture* m_texture;
    float m_scale_x = 1.f, m_scale_y = 1.f;
};

class RectangleShape {};

// records the last sprite's pixels and the last vertex array instead of drawing them
class RenderTarget {
public:
    Vector2u getSize() const { return m_size; }
    void draw(const Sprite& sprite) {
        m_pixels = sprite.m_texture->m_pixels;
        m_scale_x = sprite.m_scale_x;
        m_scale_y = sprite.m_scale_y;
    }
    void draw(const Vertex* vertices, std::size_t count, PrimitiveType type) {
        m_vertices.assign(vertices, vertices + count);
        m_type = type;
    }
    Vector2u m_size{800u, 800u};
    std::vector<Uint8> m_pixels;
    std::vector<Vertex> m_vertices;
    float m_scale_x = 0.f, m_scale_y = 0.f;
    PrimitiveType m_type = Points;
};

}  // namespace sf
