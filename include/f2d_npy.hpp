// f2d_npy.hpp -- minimal NumPy .npy (v1.0) reader/writer for 2-D float32 fields.
// The field exchange format of the headless driver: what the reference keeps
// only in host grid<float> objects (src/simulation.hpp:64-76) can be dumped, inspected with numpy and
// loaded back as an initial state.  Header-only, no dependencies.
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace f2d_npy {

inline void save(const std::string& path, const float* data, size_t rows, size_t cols) {
    std::string dict = "{'descr': '<f4', 'fortran_order': False, 'shape': (" + std::to_string(rows) + ", " +
                       std::to_string(cols) + "), }";
    size_t unpadded = 10 + dict.size() + 1;  // magic(6) + version(2) + len(2) + dict + '\n'
    size_t pad = (64 - unpadded % 64) % 64;
    dict.append(pad, ' ');
    dict.push_back('\n');
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("cannot open " + path);
    const unsigned char magic[8] = {0x93, 'N', 'U', 'M', 'P', 'Y', 1, 0};
    const uint16_t hlen = static_cast<uint16_t>(dict.size());
    bool ok = std::fwrite(magic, 1, 8, f) == 8 && std::fwrite(&hlen, 2, 1, f) == 1 &&
              std::fwrite(dict.data(), 1, dict.size(), f) == dict.size() &&
              std::fwrite(data, sizeof(float), rows * cols, f) == rows * cols;
    std::fclose(f);
    if (!ok) throw std::runtime_error("short write to " + path);
}

// Loads a C-ordered little-endian float32 2-D array; returns {rows, cols}.
inline void load(const std::string& path, std::vector<float>& out, size_t& rows, size_t& cols) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("cannot open " + path);
    unsigned char head[10];
    if (std::fread(head, 1, 10, f) != 10 || std::memcmp(head, "\x93NUMPY", 6) != 0 || head[6] != 1) {
        std::fclose(f);
        throw std::runtime_error(path + ": not a version-1 .npy file");
    }
    const size_t hlen = head[8] | (static_cast<size_t>(head[9]) << 8);
    std::string dict(hlen, '\0');
    if (std::fread(&dict[0], 1, hlen, f) != hlen) {
        std::fclose(f);
        throw std::runtime_error(path + ": truncated header");
    }
    if (dict.find("'<f4'") == std::string::npos || dict.find("'fortran_order': False") == std::string::npos) {
        std::fclose(f);
        throw std::runtime_error(path + ": need C-ordered little-endian float32");
    }
    const size_t p = dict.find("'shape': (");
    unsigned long r = 0, c = 0;
    if (p == std::string::npos || std::sscanf(dict.c_str() + p, "'shape': (%lu, %lu", &r, &c) != 2) {
        std::fclose(f);
        throw std::runtime_error(path + ": need a 2-D shape");
    }
    rows = r;
    cols = c;
    out.resize(rows * cols);
    const bool ok = std::fread(out.data(), sizeof(float), out.size(), f) == out.size();
    std::fclose(f);
    if (!ok) throw std::runtime_error(path + ": truncated data");
}

}  // namespace f2d_npy
