// ORACLE — TEST INFRASTRUCTURE ONLY (see expr.hpp header).
//
// Minimal CSC container standing in for Eigen::SparseMatrix<double, ColMajor,
// int> (SURVEY Appendix D lists the Eigen surface the path uses). Eigen itself
// is an un-vendored dependency of the reference (CMakeLists.txt:75-89, pinned
// c92d9c37) and is absent from this image, so these are restatements of its
// published behaviour: setFromTriplets sums duplicates and keeps explicit
// zeros; products/sums produce the structural union.
#pragma once

#include <algorithm>
#include <cassert>
#include <cmath>
#include <numeric>
#include <vector>

namespace orc {

using Vec = std::vector<double>;

struct Csc {
  int rows = 0, cols = 0;
  std::vector<int> colptr;  // cols + 1
  std::vector<int> rowidx;  // nnz, sorted within each column
  std::vector<double> val;  // nnz

  Csc() : colptr{0} {}
  Csc(int r, int c) : rows{r}, cols{c}, colptr(c + 1, 0) {}
  int nnz() const { return static_cast<int>(rowidx.size()); }

  /// Eigen setFromTriplets semantics: any order, duplicates summed (in input
  /// order), explicit zeros kept.
  template <class T>
  static Csc from_triplets(int rows, int cols, const std::vector<T>& trips) {
    Csc m{rows, cols};
    std::vector<int> order(trips.size());
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
      if (trips[a].col() != trips[b].col()) {
        return trips[a].col() < trips[b].col();
      }
      return trips[a].row() < trips[b].row();
    });
    int prev_r = -1, prev_c = -1;
    for (int k : order) {
      const auto& t = trips[k];
      assert(t.row() >= 0 && t.row() < rows && t.col() >= 0 && t.col() < cols);
      if (t.row() == prev_r && t.col() == prev_c) {
        m.val.back() += t.value();
      } else {
        m.rowidx.push_back(t.row());
        m.val.push_back(t.value());
        ++m.colptr[t.col() + 1];
        prev_r = t.row();
        prev_c = t.col();
      }
    }
    for (int c = 0; c < cols; ++c) m.colptr[c + 1] += m.colptr[c];
    return m;
  }

  Csc transpose() const {
    Csc t{cols, rows};
    t.rowidx.resize(nnz());
    t.val.resize(nnz());
    for (int k = 0; k < nnz(); ++k) ++t.colptr[rowidx[k] + 1];
    for (int r = 0; r < rows; ++r) t.colptr[r + 1] += t.colptr[r];
    std::vector<int> next(t.colptr.begin(), t.colptr.end() - 1);
    for (int c = 0; c < cols; ++c) {
      for (int k = colptr[c]; k < colptr[c + 1]; ++k) {
        int q = next[rowidx[k]]++;
        t.rowidx[q] = c;
        t.val[q] = val[k];
      }
    }
    return t;
  }

  /// Keep only row >= col entries (triangularView<Lower>).
  Csc lower() const {
    Csc m{rows, cols};
    for (int c = 0; c < cols; ++c) {
      for (int k = colptr[c]; k < colptr[c + 1]; ++k) {
        if (rowidx[k] >= c) {
          m.rowidx.push_back(rowidx[k]);
          m.val.push_back(val[k]);
        }
      }
      m.colptr[c + 1] = m.nnz();
    }
    return m;
  }

  /// y = A x
  Vec mul(const Vec& x) const {
    Vec y(rows, 0.0);
    for (int c = 0; c < cols; ++c) {
      for (int k = colptr[c]; k < colptr[c + 1]; ++k) {
        y[rowidx[k]] += val[k] * x[c];
      }
    }
    return y;
  }
  /// y = Aᵀ x
  Vec mul_t(const Vec& x) const {
    Vec y(cols, 0.0);
    for (int c = 0; c < cols; ++c) {
      double acc = 0.0;
      for (int k = colptr[c]; k < colptr[c + 1]; ++k) {
        acc += val[k] * x[rowidx[k]];
      }
      y[c] = acc;
    }
    return y;
  }

  /// diag(d) * A
  Csc scale_rows(const Vec& d) const {
    Csc m = *this;
    for (int k = 0; k < nnz(); ++k) m.val[k] = d[rowidx[k]] * val[k];
    return m;
  }
  /// A * diag(d)
  Csc scale_cols(const Vec& d) const {
    Csc m = *this;
    for (int c = 0; c < cols; ++c) {
      for (int k = colptr[c]; k < colptr[c + 1]; ++k) m.val[k] = val[k] * d[c];
    }
    return m;
  }
  Csc scaled(double a) const {
    Csc m = *this;
    for (auto& v : m.val) v *= a;
    return m;
  }
  bool all_finite() const {
    return std::all_of(val.begin(), val.end(),
                       [](double v) { return std::isfinite(v); });
  }
  /// Per-row ∞-norms (sparse_inf_norms.hpp:14-32).
  Vec row_inf_norms() const {
    Vec n(rows, 0.0);
    for (int k = 0; k < nnz(); ++k) {
      n[rowidx[k]] = std::max(n[rowidx[k]], std::abs(val[k]));
    }
    return n;
  }
  Csc resized(int r, int c) const {  // only growing is used
    assert(r >= rows && c >= cols);
    Csc m = *this;
    m.rows = r;
    m.cols = c;
    m.colptr.resize(c + 1, nnz());
    return m;
  }
  double coeff(int r, int c) const {
    for (int k = colptr[c]; k < colptr[c + 1]; ++k) {
      if (rowidx[k] == r) return val[k];
    }
    return 0.0;
  }
};

/// Structural-union sum a + b.
inline Csc add(const Csc& a, const Csc& b) {
  assert(a.rows == b.rows && a.cols == b.cols);
  Csc m{a.rows, a.cols};
  for (int c = 0; c < a.cols; ++c) {
    int i = a.colptr[c], ie = a.colptr[c + 1];
    int j = b.colptr[c], je = b.colptr[c + 1];
    while (i < ie || j < je) {
      if (j >= je || (i < ie && a.rowidx[i] < b.rowidx[j])) {
        m.rowidx.push_back(a.rowidx[i]);
        m.val.push_back(a.val[i++]);
      } else if (i >= ie || b.rowidx[j] < a.rowidx[i]) {
        m.rowidx.push_back(b.rowidx[j]);
        m.val.push_back(b.val[j++]);
      } else {
        m.rowidx.push_back(a.rowidx[i]);
        m.val.push_back(a.val[i++] + b.val[j++]);
      }
    }
    m.colptr[c + 1] = m.nnz();
  }
  return m;
}

/// Sparse product a * b (column-wise gather, rows sorted).
inline Csc matmul(const Csc& a, const Csc& b) {
  assert(a.cols == b.rows);
  Csc m{a.rows, b.cols};
  std::vector<double> acc(a.rows, 0.0);
  std::vector<int> mark(a.rows, -1);
  std::vector<int> pattern;
  for (int c = 0; c < b.cols; ++c) {
    pattern.clear();
    for (int k = b.colptr[c]; k < b.colptr[c + 1]; ++k) {
      int j = b.rowidx[k];
      double bv = b.val[k];
      for (int p = a.colptr[j]; p < a.colptr[j + 1]; ++p) {
        int r = a.rowidx[p];
        if (mark[r] != c) {
          mark[r] = c;
          acc[r] = 0.0;
          pattern.push_back(r);
        }
        acc[r] += a.val[p] * bv;
      }
    }
    std::sort(pattern.begin(), pattern.end());
    for (int r : pattern) {
      m.rowidx.push_back(r);
      m.val.push_back(acc[r]);
    }
    m.colptr[c + 1] = m.nnz();
  }
  return m;
}

/// Sparse diagonal from a dense vector (keeps explicit zeros, like
/// SparseMatrix{vec.asDiagonal()}).
inline Csc diag(const Vec& d) {
  int n = static_cast<int>(d.size());
  Csc m{n, n};
  m.rowidx.resize(n);
  m.val = d;
  for (int i = 0; i < n; ++i) {
    m.rowidx[i] = i;
    m.colptr[i + 1] = i + 1;
  }
  return m;
}

/// Stack blocks vertically: [a; b] (append_as_triplets.hpp:25-48 produces
/// exactly this column-major interleave).
inline Csc vstack(const Csc& a, const Csc& b) {
  assert(a.cols == b.cols);
  Csc m{a.rows + b.rows, a.cols};
  for (int c = 0; c < a.cols; ++c) {
    for (int k = a.colptr[c]; k < a.colptr[c + 1]; ++k) {
      m.rowidx.push_back(a.rowidx[k]);
      m.val.push_back(a.val[k]);
    }
    for (int k = b.colptr[c]; k < b.colptr[c + 1]; ++k) {
      m.rowidx.push_back(a.rows + b.rowidx[k]);
      m.val.push_back(b.val[k]);
    }
    m.colptr[c + 1] = m.nnz();
  }
  return m;
}

// Dense vector helpers (Eigen lpNorm<1>, lpNorm<Infinity>, norm, dot).
inline double norm1(const Vec& v) {
  double s = 0.0;
  for (double x : v) s += std::abs(x);
  return s;
}
inline double norm_inf(const Vec& v) {
  double s = 0.0;
  for (double x : v) s = std::max(s, std::abs(x));
  return s;
}
inline double norm2(const Vec& v) {
  double s = 0.0;
  for (double x : v) s += x * x;
  return std::sqrt(s);
}
inline double dot(const Vec& a, const Vec& b) {
  double s = 0.0;
  for (size_t i = 0; i < a.size(); ++i) s += a[i] * b[i];
  return s;
}
inline bool all_finite(const Vec& v) {
  return std::all_of(v.begin(), v.end(),
                     [](double x) { return std::isfinite(x); });
}
inline Vec sub(const Vec& a, const Vec& b) {
  Vec r(a.size());
  for (size_t i = 0; i < a.size(); ++i) r[i] = a[i] - b[i];
  return r;
}
inline Vec axpy(const Vec& x, double a, const Vec& p) {  // x + a p
  Vec r(x.size());
  for (size_t i = 0; i < x.size(); ++i) r[i] = x[i] + a * p[i];
  return r;
}

}  // namespace orc
