// ORACLE — TEST INFRASTRUCTURE ONLY (see expr.hpp header).
//
// PARITY UNPINNED at this boundary: the arithmetic below lives in Eigen
// (Eigen::SimplicialLDLT<SparseMatrix<double>>, default Lower + AMDOrdering,
// and Eigen::LDLT<MatrixXd>), an un-vendored dependency of the reference
// pinned to git c92d9c379dd034d3e7ddb7cdd7ba0add3bf1c747
// (/root/reference/CMakeLists.txt:75-89) that is absent from this image, and
// no reference test touches the factor. What follows restates the
// published algorithms Eigen implements:
//   * approximate minimum degree ordering (Amestoy, Davis, Duff 1996/2004) on
//     the full symmetric pattern, diagonal kept, "dense or no structural
//     diagonal" rows ordered last, assembly-tree postorder;
//   * elimination tree + column counts by row-subtree traversal, and the
//     up-looking sparse LDLᵀ of Davis' LDL (no pivoting; fails iff a pivot is
//     exactly 0);
//   * dense LDLᵀ with symmetric diagonal pivoting (largest |diagonal|).
// Call sites being replaced: solver/util/sparse_regularized_ldlt.hpp:70,74,105,
// 78,109,160; dense_regularized_ldlt.hpp:60,90,167;
// lagrange_multiplier_estimate.hpp:42-44,107-108.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <numeric>
#include <vector>

#include "sparse.hpp"

namespace orc {

/// Approximate minimum degree on a symmetric pattern (both triangles +
/// diagonal present). Returns p with p[k] = original index of the k-th pivot.
inline std::vector<int> amd_order(int n, const std::vector<int>& Ap,
                                  const std::vector<int>& Ai) {
  if (n == 0) return {};
  auto flip = [](int i) { return -i - 2; };

  int dense = std::max(16, static_cast<int>(10 * std::sqrt(double(n))));
  dense = std::min(n - 2, dense);

  int cnz = Ap[n];
  std::vector<int> Cp(Ap.begin(), Ap.end());
  int nzmax = cnz + cnz / 5 + 2 * n;
  std::vector<int> Ci(nzmax);
  std::copy(Ai.begin(), Ai.begin() + cnz, Ci.begin());

  std::vector<int> len(n + 1), nv(n + 1), next(n + 1), head(n + 1),
      elen(n + 1), degree(n + 1), w(n + 1), hhead(n + 1), last(n + 1);

  for (int k = 0; k < n; ++k) len[k] = Cp[k + 1] - Cp[k];
  len[n] = 0;
  for (int i = 0; i <= n; ++i) {
    head[i] = -1;
    last[i] = -1;
    next[i] = -1;
    hhead[i] = -1;
    nv[i] = 1;
    w[i] = 1;
    elen[i] = 0;
    degree[i] = len[i];
  }
  auto wclear = [&](int mark, int lemax) {
    if (mark < 2 || (mark + lemax < 0)) {
      for (int k = 0; k < n; ++k) {
        if (w[k] != 0) w[k] = 1;
      }
      mark = 2;
    }
    return mark;
  };
  int mark = wclear(0, 0);
  int nel = 0, mindeg = 0, lemax = 0;

  // Degree lists. A row whose only entry is its own diagonal is "empty"; a
  // dense row, or one lacking a structural diagonal, is deferred to the end.
  for (int i = 0; i < n; ++i) {
    bool has_diag = false;
    for (int p = Cp[i]; p < Cp[i + 1]; ++p) {
      if (Ci[p] == i) {
        has_diag = true;
        break;
      }
    }
    int d = degree[i];
    if (d == 1 && has_diag) {
      elen[i] = -2;
      ++nel;
      Cp[i] = -1;
      w[i] = 0;
    } else if (d > dense || !has_diag) {
      nv[i] = 0;
      elen[i] = -1;
      ++nel;
      Cp[i] = flip(n);
      ++nv[n];
    } else {
      if (head[d] != -1) last[head[d]] = i;
      next[i] = head[d];
      head[d] = i;
    }
  }
  elen[n] = -2;
  Cp[n] = -1;
  w[n] = 0;

  while (nel < n) {
    // Pivot of minimum approximate degree.
    int k = -1;
    for (; mindeg < n && (k = head[mindeg]) == -1; ++mindeg) {
    }
    if (next[k] != -1) last[next[k]] = -1;
    head[mindeg] = next[k];
    int elenk = elen[k];
    int nvk = nv[k];
    nel += nvk;

    // Compact the quotient graph when the elbow room runs out.
    if (elenk > 0 && cnz + mindeg >= nzmax) {
      for (int j = 0; j < n; ++j) {
        int p = Cp[j];
        if (p >= 0) {
          Cp[j] = Ci[p];
          Ci[p] = flip(j);
        }
      }
      int q = 0;
      for (int p = 0; p < cnz;) {
        int j = flip(Ci[p++]);
        if (j >= 0) {
          Ci[q] = Cp[j];
          Cp[j] = q++;
          for (int k3 = 0; k3 < len[j] - 1; ++k3) Ci[q++] = Ci[p++];
        }
      }
      cnz = q;
    }

    // New element Lk = union of k's variables and its elements' variables.
    int dk = 0;
    nv[k] = -nvk;
    int p = Cp[k];
    int pk1 = (elenk == 0) ? p : cnz;
    int pk2 = pk1;
    for (int k1 = 1; k1 <= elenk + 1; ++k1) {
      int e, pj, ln;
      if (k1 > elenk) {
        e = k;
        pj = p;
        ln = len[k] - elenk;
      } else {
        e = Ci[p++];
        pj = Cp[e];
        ln = len[e];
      }
      for (int k2 = 1; k2 <= ln; ++k2) {
        int i = Ci[pj++];
        int nvi = nv[i];
        if (nvi <= 0) continue;
        dk += nvi;
        nv[i] = -nvi;
        Ci[pk2++] = i;
        if (next[i] != -1) last[next[i]] = last[i];
        if (last[i] != -1) {
          next[last[i]] = next[i];
        } else {
          head[degree[i]] = next[i];
        }
      }
      if (e != k) {
        Cp[e] = flip(k);
        w[e] = 0;
      }
    }
    if (elenk != 0) cnz = pk2;
    degree[k] = dk;
    Cp[k] = pk1;
    len[k] = pk2 - pk1;
    elen[k] = -2;

    // Scan 1: |Le \ Lk| for every element e adjacent to a variable of Lk.
    mark = wclear(mark, lemax);
    for (int pk = pk1; pk < pk2; ++pk) {
      int i = Ci[pk];
      int eln = elen[i];
      if (eln <= 0) continue;
      int nvi = -nv[i];
      int wnvi = mark - nvi;
      for (int q = Cp[i]; q <= Cp[i] + eln - 1; ++q) {
        int e = Ci[q];
        if (w[e] >= mark) {
          w[e] -= nvi;
        } else if (w[e] != 0) {
          w[e] = degree[e] + wnvi;
        }
      }
    }

    // Scan 2: approximate degree update, aggressive absorption, hashing.
    for (int pk = pk1; pk < pk2; ++pk) {
      int i = Ci[pk];
      int p1 = Cp[i];
      int p2 = p1 + elen[i] - 1;
      int pn = p1;
      int d = 0;
      unsigned h = 0;
      for (int q = p1; q <= p2; ++q) {
        int e = Ci[q];
        if (w[e] != 0) {
          int dext = w[e] - mark;
          if (dext > 0) {
            d += dext;
            Ci[pn++] = e;
            h += static_cast<unsigned>(e);
          } else {
            Cp[e] = flip(k);
            w[e] = 0;
          }
        }
      }
      elen[i] = pn - p1 + 1;
      int p3 = pn;
      int p4 = p1 + len[i];
      for (int q = p2 + 1; q < p4; ++q) {
        int j = Ci[q];
        int nvj = nv[j];
        if (nvj <= 0) continue;
        d += nvj;
        Ci[pn++] = j;
        h += static_cast<unsigned>(j);
      }
      if (d == 0) {
        // Mass elimination: i is indistinguishable from k.
        Cp[i] = flip(k);
        int nvi = -nv[i];
        dk -= nvi;
        nvk += nvi;
        nel += nvi;
        nv[i] = 0;
        elen[i] = -1;
      } else {
        degree[i] = std::min(degree[i], d);
        Ci[pn] = Ci[p3];
        Ci[p3] = Ci[p1];
        Ci[p1] = k;
        len[i] = pn - p1 + 1;
        h %= static_cast<unsigned>(n);
        next[i] = hhead[h];
        hhead[h] = i;
        last[i] = static_cast<int>(h);
      }
    }
    degree[k] = dk;
    lemax = std::max(lemax, dk);
    mark = wclear(mark + lemax, lemax);

    // Supervariable detection among the hashed members of Lk.
    for (int pk = pk1; pk < pk2; ++pk) {
      int i = Ci[pk];
      if (nv[i] >= 0) continue;
      int h = last[i];
      i = hhead[h];
      hhead[h] = -1;
      for (; i != -1 && next[i] != -1; i = next[i], ++mark) {
        int ln = len[i];
        int eln = elen[i];
        for (int q = Cp[i] + 1; q <= Cp[i] + ln - 1; ++q) w[Ci[q]] = mark;
        int jlast = i;
        for (int j = next[i]; j != -1;) {
          bool ok = (len[j] == ln) && (elen[j] == eln);
          for (int q = Cp[j] + 1; ok && q <= Cp[j] + ln - 1; ++q) {
            if (w[Ci[q]] != mark) ok = false;
          }
          if (ok) {
            Cp[j] = flip(i);
            nv[i] += nv[j];
            nv[j] = 0;
            elen[j] = -1;
            j = next[j];
            next[jlast] = j;
          } else {
            jlast = j;
            j = next[j];
          }
        }
      }
    }

    // Finalise Lk: restore nv, recompute external degrees, relink lists.
    int pf = pk1;
    for (int pk = pk1; pk < pk2; ++pk) {
      int i = Ci[pk];
      int nvi = -nv[i];
      if (nvi <= 0) continue;
      nv[i] = nvi;
      int d = degree[i] + dk - nvi;
      d = std::min(d, n - nel - nvi);
      if (head[d] != -1) last[head[d]] = i;
      next[i] = head[d];
      last[i] = -1;
      head[d] = i;
      mindeg = std::min(mindeg, d);
      degree[i] = d;
      Ci[pf++] = i;
    }
    nv[k] = nvk;
    if ((len[k] = pf - pk1) == 0) {
      Cp[k] = -1;
      w[k] = 0;
    }
    if (elenk != 0) cnz = pf;
  }

  // Postorder the assembly tree.
  for (int i = 0; i < n; ++i) Cp[i] = flip(Cp[i]);
  for (int j = 0; j <= n; ++j) head[j] = -1;
  for (int j = n; j >= 0; --j) {
    if (nv[j] > 0) continue;
    next[j] = head[Cp[j]];
    head[Cp[j]] = j;
  }
  for (int e = n; e >= 0; --e) {
    if (nv[e] <= 0) continue;
    if (Cp[e] != -1) {
      next[e] = head[Cp[e]];
      head[Cp[e]] = e;
    }
  }
  std::vector<int> post(n + 1);
  int k = 0;
  for (int i = 0; i <= n; ++i) {
    if (Cp[i] != -1) continue;
    // depth-first traversal from root i
    int top = 0;
    w[0] = i;
    while (top >= 0) {
      int pnode = w[top];
      int child = head[pnode];
      if (child == -1) {
        --top;
        post[k++] = pnode;
      } else {
        head[pnode] = next[child];
        w[++top] = child;
      }
    }
  }
  post.resize(n);
  return post;
}

/// Full symmetric pattern (both triangles, diagonal kept) of a matrix given by
/// its lower triangle.
inline void symmetrize_pattern(const Csc& lower, std::vector<int>& Ap,
                               std::vector<int>& Ai) {
  int n = lower.cols;
  std::vector<int> cnt(n, 0);
  for (int c = 0; c < n; ++c) {
    for (int k = lower.colptr[c]; k < lower.colptr[c + 1]; ++k) {
      int r = lower.rowidx[k];
      ++cnt[c];
      if (r != c) ++cnt[r];
    }
  }
  Ap.assign(n + 1, 0);
  for (int c = 0; c < n; ++c) Ap[c + 1] = Ap[c] + cnt[c];
  Ai.assign(Ap[n], 0);
  std::vector<int> nxt(Ap.begin(), Ap.end() - 1);
  // Two passes keep each column sorted: first the upper part (rows < c) in
  // increasing order, then the lower part.
  for (int c = 0; c < n; ++c) {
    for (int k = lower.colptr[c]; k < lower.colptr[c + 1]; ++k) {
      int r = lower.rowidx[k];
      if (r != c) Ai[nxt[r]++] = c;  // entry (c, r) in column r, row c < r
    }
  }
  for (int c = 0; c < n; ++c) {
    for (int k = lower.colptr[c]; k < lower.colptr[c + 1]; ++k) {
      Ai[nxt[c]++] = lower.rowidx[k];
    }
  }
}

/// Simplicial sparse LDLᵀ with a fixed symbolic analysis.
class SimplicialLDLT {
 public:
  enum class Ordering { AMD, NATURAL, CUSTOM };

  /// p[k] = original index eliminated k-th. Must be set before analyze().
  void set_custom_permutation(std::vector<int> p) {
    m_ordering = Ordering::CUSTOM;
    m_p = std::move(p);
  }
  void set_ordering(Ordering o) { m_ordering = o; }

  /// `lower`: lower triangle (CSC) of the symmetric matrix.
  void analyze(const Csc& lower) {
    n = lower.cols;
    if (m_ordering == Ordering::AMD) {
      std::vector<int> Ap, Ai;
      symmetrize_pattern(lower, Ap, Ai);
      m_p = amd_order(n, Ap, Ai);
    } else if (m_ordering == Ordering::NATURAL) {
      m_p.resize(n);
      std::iota(m_p.begin(), m_p.end(), 0);
    }
    m_pinv.assign(n, 0);
    for (int k = 0; k < n; ++k) m_pinv[m_p[k]] = k;

    // Permuted upper-triangular pattern B = A(p,p), column k holds rows <= k,
    // with a map back to the entries of `lower`.
    std::vector<int> cnt(n, 0);
    for (int c = 0; c < n; ++c) {
      for (int k = lower.colptr[c]; k < lower.colptr[c + 1]; ++k) {
        int a = m_pinv[lower.rowidx[k]], b = m_pinv[c];
        ++cnt[std::max(a, b)];
      }
    }
    m_up.assign(n + 1, 0);
    for (int c = 0; c < n; ++c) m_up[c + 1] = m_up[c] + cnt[c];
    m_ui.assign(m_up[n], 0);
    m_usrc.assign(m_up[n], 0);
    {
      // Fill by increasing row so columns come out sorted.
      std::vector<std::vector<std::pair<int, int>>> byrow(n);
      for (int c = 0; c < n; ++c) {
        for (int k = lower.colptr[c]; k < lower.colptr[c + 1]; ++k) {
          int a = m_pinv[lower.rowidx[k]], b = m_pinv[c];
          byrow[std::min(a, b)].emplace_back(std::max(a, b), k);
        }
      }
      std::vector<int> nxt(m_up.begin(), m_up.end() - 1);
      for (int r = 0; r < n; ++r) {
        for (auto [c, src] : byrow[r]) {
          m_ui[nxt[c]] = r;
          m_usrc[nxt[c]++] = src;
        }
      }
    }

    // Elimination tree and column counts (row-subtree traversal).
    m_parent.assign(n, -1);
    m_lnz.assign(n, 0);
    std::vector<int> tags(n);
    for (int k = 0; k < n; ++k) {
      tags[k] = k;
      for (int q = m_up[k]; q < m_up[k + 1]; ++q) {
        int i = m_ui[q];
        if (i < k) {
          for (; tags[i] != k; i = m_parent[i]) {
            if (m_parent[i] == -1) m_parent[i] = k;
            ++m_lnz[i];
            tags[i] = k;
          }
        }
      }
    }
    m_lp.assign(n + 1, 0);
    for (int k = 0; k < n; ++k) m_lp[k + 1] = m_lp[k] + m_lnz[k];
    m_li.assign(m_lp[n], 0);
    m_lx.assign(m_lp[n], 0.0);
    m_d.assign(n, 0.0);
    m_analyzed = true;
  }

  /// Numeric up-looking LDLᵀ. Returns false iff a pivot is exactly zero
  /// (Eigen: info() == NumericalIssue).
  bool factorize(const Csc& lower) {
    std::vector<double> y(n, 0.0);
    std::vector<int> pattern(n), tags(n), nzcol(n, 0);
    bool ok = true;
    for (int k = 0; k < n; ++k) {
      y[k] = 0.0;
      int top = n;
      tags[k] = k;
      nzcol[k] = 0;
      for (int q = m_up[k]; q < m_up[k + 1]; ++q) {
        int i = m_ui[q];
        if (i <= k) {
          y[i] += lower.val[m_usrc[q]];
          int len = 0;
          for (; tags[i] != k; i = m_parent[i]) {
            pattern[len++] = i;
            tags[i] = k;
          }
          while (len > 0) pattern[--top] = pattern[--len];
        }
      }
      double d = y[k];
      y[k] = 0.0;
      for (; top < n; ++top) {
        int i = pattern[top];
        double yi = y[i];
        y[i] = 0.0;
        double l_ki = yi / m_d[i];
        int p2 = m_lp[i] + nzcol[i];
        for (int p = m_lp[i]; p < p2; ++p) y[m_li[p]] -= m_lx[p] * yi;
        d -= l_ki * yi;
        m_li[p2] = k;
        m_lx[p2] = l_ki;
        ++nzcol[i];
      }
      m_d[k] = d;
      if (d == 0.0) {
        ok = false;
        break;
      }
    }
    m_ok = ok;
    return ok;
  }

  const std::vector<double>& vectorD() const { return m_d; }
  bool ok() const { return m_ok; }
  const std::vector<int>& permutation() const { return m_p; }
  int nnz_l() const { return m_lp.empty() ? 0 : m_lp[n]; }
  const std::vector<int>& parent() const { return m_parent; }
  bool analyzed() const { return m_analyzed; }

  Vec solve(const Vec& b) const {
    Vec x(n);
    for (int k = 0; k < n; ++k) x[k] = b[m_p[k]];
    for (int j = 0; j < n; ++j) {
      double xj = x[j];
      for (int p = m_lp[j]; p < m_lp[j + 1]; ++p) x[m_li[p]] -= m_lx[p] * xj;
    }
    for (int j = 0; j < n; ++j) x[j] /= m_d[j];
    for (int j = n - 1; j >= 0; --j) {
      double acc = x[j];
      for (int p = m_lp[j]; p < m_lp[j + 1]; ++p) acc -= m_lx[p] * x[m_li[p]];
      x[j] = acc;
    }
    Vec out(n);
    for (int k = 0; k < n; ++k) out[m_p[k]] = x[k];
    return out;
  }

  int n = 0;

 private:
  Ordering m_ordering = Ordering::AMD;
  bool m_analyzed = false, m_ok = false;
  std::vector<int> m_p, m_pinv;
  std::vector<int> m_up, m_ui, m_usrc;  // permuted upper pattern + source map
  std::vector<int> m_parent, m_lnz, m_lp, m_li;
  std::vector<double> m_lx, m_d;
};

/// Dense LDLᵀ with symmetric diagonal pivoting (Eigen::LDLT semantics: at step
/// k the largest remaining |diagonal| is swapped in; D is diagonal).
class DenseLDLT {
 public:
  /// `a`: full dense symmetric matrix, row-major n×n; only the lower triangle
  /// is read.
  bool compute(std::vector<double> a, int n_) {
    n = n_;
    m = std::move(a);
    m_tr.resize(n);
    bool ok = true;
    std::vector<double> temp(n);
    auto at = [&](int r, int c) -> double& { return m[r * n + c]; };
    for (int k = 0; k < n; ++k) {
      int piv = k;
      double big = std::abs(at(k, k));
      for (int i = k + 1; i < n; ++i) {
        if (std::abs(at(i, i)) > big) {
          big = std::abs(at(i, i));
          piv = i;
        }
      }
      m_tr[k] = piv;
      if (piv != k) {
        // symmetric swap of rows/cols k and piv on the lower triangle
        int s = n - piv - 1;
        for (int c = 0; c < k; ++c) std::swap(at(k, c), at(piv, c));
        for (int r = 0; r < s; ++r) {
          std::swap(at(piv + 1 + r, k), at(piv + 1 + r, piv));
        }
        std::swap(at(k, k), at(piv, piv));
        for (int i = k + 1; i < piv; ++i) std::swap(at(i, k), at(piv, i));
      }
      int rs = n - k - 1;
      if (k > 0) {
        for (int c = 0; c < k; ++c) temp[c] = at(c, c) * at(k, c);
        double acc = 0.0;
        for (int c = 0; c < k; ++c) acc += at(k, c) * temp[c];
        at(k, k) -= acc;
        for (int r = 0; r < rs; ++r) {
          double a2 = 0.0;
          for (int c = 0; c < k; ++c) a2 += at(k + 1 + r, c) * temp[c];
          at(k + 1 + r, k) -= a2;
        }
      }
      double akk = at(k, k);
      bool valid = std::abs(akk) > 0.0;
      if (k == 0 && !valid) {
        // The whole diagonal is zero: nothing to pivot on.
        for (int j = 0; j < n; ++j) m_tr[j] = j;
        bool all_zero = true;
        for (int r = 1; r < n && all_zero; ++r) {
          for (int c = 0; c < r; ++c) {
            if (at(r, c) != 0.0) {
              all_zero = false;
              break;
            }
          }
        }
        ok = all_zero;
        break;
      }
      if (rs > 0 && valid) {
        for (int r = 0; r < rs; ++r) at(k + 1 + r, k) /= akk;
      } else if (rs > 0) {
        for (int r = 0; r < rs; ++r) ok = ok && (at(k + 1 + r, k) == 0.0);
      }
    }
    m_ok = ok;
    return ok;
  }

  std::vector<double> vectorD() const {
    std::vector<double> d(n);
    for (int i = 0; i < n; ++i) d[i] = m[i * n + i];
    return d;
  }

  Vec solve(const Vec& b) const {
    Vec x = b;
    for (int k = 0; k < n; ++k) std::swap(x[k], x[m_tr[k]]);
    for (int i = 0; i < n; ++i) {
      double acc = x[i];
      for (int c = 0; c < i; ++c) acc -= m[i * n + c] * x[c];
      x[i] = acc;
    }
    // Pseudo-inverse of D with Eigen's tolerance (smallest normal number).
    const double tol = std::numeric_limits<double>::min();
    for (int i = 0; i < n; ++i) {
      double d = m[i * n + i];
      if (std::abs(d) > tol) {
        x[i] /= d;
      } else {
        x[i] = 0.0;
      }
    }
    for (int i = n - 1; i >= 0; --i) {
      double acc = x[i];
      for (int r = i + 1; r < n; ++r) acc -= m[r * n + i] * x[r];
      x[i] = acc;
    }
    for (int k = n - 1; k >= 0; --k) std::swap(x[k], x[m_tr[k]]);
    return x;
  }

  bool ok() const { return m_ok; }
  int n = 0;

 private:
  std::vector<double> m;
  std::vector<int> m_tr;
  bool m_ok = false;
};

}  // namespace orc
