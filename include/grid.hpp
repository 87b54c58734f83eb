// grid.hpp -- host field container with the API of the reference's grid<T> (src/grid.hpp:8-66):
// ctor (rows, cols, initial_value), rows(), cols(), data(), operator()(i, j) row-major, begin/end.
// Written fresh for this repository; two deliberate differences from the reference:
//   * cols() returns the column count (the reference returns m_rows, src/grid.hpp:20-22);
//   * <cstddef> is included (the reference relies on a transitive include for size_t).
// Code written against the reference's grid<T> compiles unchanged against this one.
#pragma once

#include <cstddef>
#include <vector>

template <typename T>
class grid {
public:
    grid(std::size_t rows, std::size_t cols, T initial_value)
        : rows_(rows), cols_(cols), cells_(rows * cols, initial_value) {}

    std::size_t rows() const { return rows_; }
    std::size_t cols() const { return cols_; }

    const T* data() const { return cells_.data(); }
    T* data() { return cells_.data(); }

    // value at row i, column j (row-major, pitch == cols)
    T operator()(const std::size_t i, const std::size_t j) const { return cells_[i * cols_ + j]; }
    T& operator()(const std::size_t i, const std::size_t j) { return cells_[i * cols_ + j]; }

    auto begin() { return cells_.begin(); }
    auto end() { return cells_.end(); }
    auto cbegin() const { return cells_.cbegin(); }
    auto cend() const { return cells_.cend(); }

private:
    std::size_t rows_;
    std::size_t cols_;
    std::vector<T> cells_;
};
