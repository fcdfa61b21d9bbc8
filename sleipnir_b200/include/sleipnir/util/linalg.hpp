// Minimal host containers shaped like the Eigen types that appear in the
// reference's public solver API (IterationInfo, callbacks): a dense vector, a
// dense row-major matrix and a compressed column-major sparse matrix with
// 32-bit indices (Eigen::SparseMatrix<double, ColMajor, int> layout, SURVEY
// §8b). Eigen is an un-vendored dependency of the reference and is not
// available to this build; these carry data across the API, they are not a
// linear-algebra library — the arithmetic of the path runs on the device.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <initializer_list>
#include <vector>

#include "sleipnir/util/assert.hpp"

namespace slp {

/// Dense column vector.
template <typename Scalar>
class Vector {
 public:
  Vector() = default;
  explicit Vector(int rows) : m_data(rows, Scalar(0)) {}
  Vector(int rows, Scalar fill) : m_data(rows, fill) {}
  Vector(std::initializer_list<Scalar> init) : m_data{init} {}
  explicit Vector(std::vector<Scalar> data) : m_data{std::move(data)} {}

  int rows() const { return static_cast<int>(m_data.size()); }
  int cols() const { return 1; }
  int size() const { return rows(); }
  void resize(int rows) { m_data.resize(rows); }
  Scalar& operator[](int i) { return m_data[i]; }
  const Scalar& operator[](int i) const { return m_data[i]; }
  Scalar& operator()(int i) { return m_data[i]; }
  const Scalar& operator()(int i) const { return m_data[i]; }
  Scalar& operator()(int r, int) { return m_data[r]; }
  const Scalar& operator()(int r, int) const { return m_data[r]; }
  Scalar* data() { return m_data.data(); }
  const Scalar* data() const { return m_data.data(); }
  auto begin() { return m_data.begin(); }
  auto end() { return m_data.end(); }
  auto begin() const { return m_data.begin(); }
  auto end() const { return m_data.end(); }
  Vector segment(int offset, int length) const {
    return Vector{std::vector<Scalar>(m_data.begin() + offset,
                                      m_data.begin() + offset + length)};
  }
  bool allFinite() const {
    for (const auto& v : m_data) {
      if (!std::isfinite(v)) return false;
    }
    return true;
  }
  Scalar coeff(int i) const { return m_data[i]; }
  Scalar lpNormInf() const {
    Scalar m(0);
    for (const auto& v : m_data) m = std::max(m, std::abs(v));
    return m;
  }

 private:
  std::vector<Scalar> m_data;
};

/// Dense row-major matrix of plain numbers (DSL operand and value() result).
template <typename Scalar>
class Matrix {
 public:
  Matrix() = default;
  Matrix(int rows, int cols)
      : m_rows{rows}, m_cols{cols}, m_data(size_t(rows) * cols, Scalar(0)) {}
  Matrix(std::initializer_list<std::initializer_list<Scalar>> init) {
    m_rows = static_cast<int>(init.size());
    m_cols = m_rows ? static_cast<int>(init.begin()->size()) : 0;
    for (const auto& row : init) {
      slp_assert(static_cast<int>(row.size()) == m_cols);
      m_data.insert(m_data.end(), row.begin(), row.end());
    }
  }
  int rows() const { return m_rows; }
  int cols() const { return m_cols; }
  Scalar& operator()(int r, int c) { return m_data[size_t(r) * m_cols + c]; }
  const Scalar& operator()(int r, int c) const {
    return m_data[size_t(r) * m_cols + c];
  }
  Scalar& operator[](int r, int c) { return (*this)(r, c); }
  const Scalar& operator[](int r, int c) const { return (*this)(r, c); }
  Scalar& operator()(int i) { return m_data[i]; }
  const Scalar& operator()(int i) const { return m_data[i]; }
  Scalar* data() { return m_data.data(); }
  const Scalar* data() const { return m_data.data(); }

 private:
  int m_rows = 0, m_cols = 0;
  std::vector<Scalar> m_data;
};

/// Compressed sparse column matrix, 32-bit indices, rows sorted per column.
template <typename Scalar>
class SparseMatrix {
 public:
  SparseMatrix() : m_outer{0} {}
  SparseMatrix(int rows, int cols)
      : m_rows{rows}, m_cols{cols}, m_outer(cols + 1, 0) {}
  SparseMatrix(int rows, int cols, std::vector<int32_t> outer,
               std::vector<int32_t> inner, std::vector<Scalar> values)
      : m_rows{rows}, m_cols{cols}, m_outer{std::move(outer)},
        m_inner{std::move(inner)}, m_values{std::move(values)} {}

  int rows() const { return m_rows; }
  int cols() const { return m_cols; }
  int64_t nonZeros() const { return static_cast<int64_t>(m_inner.size()); }
  const int32_t* outerIndexPtr() const { return m_outer.data(); }
  const int32_t* innerIndexPtr() const { return m_inner.data(); }
  const Scalar* valuePtr() const { return m_values.data(); }
  Scalar* valuePtr() { return m_values.data(); }
  Scalar coeff(int row, int col) const {
    for (int32_t k = m_outer[col]; k < m_outer[col + 1]; ++k) {
      if (m_inner[k] == row) return m_values[k];
    }
    return Scalar(0);
  }
  /// Eigen's setFromTriplets: duplicates are summed in triplet order, explicit
  /// zeros are kept, rows sorted within each column. keep(row, col) filters
  /// (triangularView).
  template <typename TripletRange, typename Keep>
  static SparseMatrix from_triplets(int rows, int cols,
                                    const TripletRange& triplets, Keep keep) {
    std::vector<int32_t> count(cols + 1, 0);
    for (const auto& t : triplets) {
      if (keep(t.row, t.col)) ++count[t.col + 1];
    }
    for (int c = 0; c < cols; ++c) count[c + 1] += count[c];
    std::vector<int32_t> r(count[cols]);
    std::vector<Scalar> v(count[cols]);
    std::vector<int32_t> nxt(count.begin(), count.end() - 1);
    for (const auto& t : triplets) {  // stable within a column
      if (!keep(t.row, t.col)) continue;
      r[nxt[t.col]] = t.row;
      v[nxt[t.col]++] = static_cast<Scalar>(t.value);
    }
    SparseMatrix out{rows, cols};
    out.m_outer.assign(cols + 1, 0);
    std::vector<int32_t> order;
    for (int c = 0; c < cols; ++c) {
      const int32_t b = count[c], e = count[c + 1];
      order.resize(e - b);
      for (int32_t k = 0; k < e - b; ++k) order[k] = b + k;
      std::stable_sort(order.begin(), order.end(),
                       [&](int32_t x, int32_t y) { return r[x] < r[y]; });
      for (size_t k = 0; k < order.size(); ++k) {
        const int32_t q = order[k];
        if (k > 0 && r[q] == out.m_inner.back()) {
          out.m_values.back() += v[q];
        } else {
          out.m_inner.push_back(r[q]);
          out.m_values.push_back(v[q]);
        }
      }
      out.m_outer[c + 1] = static_cast<int32_t>(out.m_inner.size());
    }
    return out;
  }

 private:
  int m_rows = 0, m_cols = 0;
  std::vector<int32_t> m_outer, m_inner;
  std::vector<Scalar> m_values;
};

}  // namespace slp
