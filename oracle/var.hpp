// ORACLE — TEST INFRASTRUCTURE ONLY (see expr.hpp header).
//
// Restatement of the front half of the reference DSL that the hot-path configs
// use: include/sleipnir/autodiff/variable.hpp (Variable, constraint factories
// :721-779, :896-1013) and variable_matrix.hpp (row-major storage :313, matrix
// product as a left fold from constant 0 :551-566, scalar*matrix = elem*scalar
// :627-637, T() :955-965, solve() 1x1..3x3 :1516-1600, gradient_tree
// :1756-1805). Templated on the expression backend `B` so the very same DSL
// code runs on the restated core (expr.hpp) and on the reference's own core
// (backend_ref.hpp).
#pragma once

#include <cassert>
#include <initializer_list>
#include <utility>
#include <vector>

namespace orc {

template <class B>
class Var {
 public:
  using Ptr = typename B::Ptr;
  Ptr expr;

  /// Default = fresh decision variable (variable.hpp:271-272).
  Var() : expr{B::make_var()} {}
  Var(double v) : expr{B::make_const(v)} {}  // NOLINT
  Var(int v) : expr{B::make_const(static_cast<double>(v))} {}  // NOLINT
  explicit Var(Ptr p) : expr{std::move(p)} {}
  static Var null() { return Var{Ptr{nullptr}}; }

  void set_value(double v) { expr->val = v; }  // variable.hpp:125-138

  /// variable.hpp:143-151: lazily sorted graph, forward sweep, root value.
  double value() {
    if (!m_graph_init) {
      m_graph = B::sort(expr);
      m_graph_init = true;
    }
    B::update(m_graph);
    return expr->val;
  }
  int type() const { return B::type_rank(expr); }

  friend Var operator*(const Var& l, const Var& r) {
    return Var{l.expr * r.expr};
  }
  friend Var operator/(const Var& l, const Var& r) {
    return Var{l.expr / r.expr};
  }
  friend Var operator+(const Var& l, const Var& r) {
    return Var{l.expr + r.expr};
  }
  friend Var operator-(const Var& l, const Var& r) {
    return Var{l.expr - r.expr};
  }
  friend Var operator-(const Var& l) { return Var{-l.expr}; }
  Var& operator+=(const Var& r) { return *this = *this + r; }
  Var& operator-=(const Var& r) { return *this = *this - r; }
  Var& operator*=(const Var& r) { return *this = *this * r; }
  Var& operator/=(const Var& r) { return *this = *this / r; }

 private:
  typename B::Graph m_graph;
  bool m_graph_init = false;
};

#define ORC_UNARY(fn) \
  template <class B>  \
  Var<B> fn(const Var<B>& x) { return Var<B>{fn(x.expr)}; }
ORC_UNARY(abs) ORC_UNARY(acos) ORC_UNARY(asin) ORC_UNARY(atan) ORC_UNARY(cbrt)
ORC_UNARY(cos) ORC_UNARY(cosh) ORC_UNARY(erf) ORC_UNARY(exp) ORC_UNARY(log)
ORC_UNARY(log10) ORC_UNARY(sign) ORC_UNARY(sin) ORC_UNARY(sinh) ORC_UNARY(sqrt)
ORC_UNARY(tan) ORC_UNARY(tanh)
#undef ORC_UNARY
#define ORC_BINARY(fn)                                                      \
  template <class B>                                                        \
  Var<B> fn(const Var<B>& a, const Var<B>& b) {                             \
    return Var<B>{fn(a.expr, b.expr)};                                      \
  }                                                                         \
  template <class B>                                                        \
  Var<B> fn(const Var<B>& a, double b) { return fn(a, Var<B>{b}); }         \
  template <class B>                                                        \
  Var<B> fn(double a, const Var<B>& b) { return fn(Var<B>{a}, b); }
ORC_BINARY(atan2) ORC_BINARY(hypot) ORC_BINARY(max) ORC_BINARY(min)
ORC_BINARY(pow)
#undef ORC_BINARY

/// hypot(x, y, z) = sqrt(x² + y² + z²) (variable.hpp:711-714).
template <class B>
Var<B> hypot3(const Var<B>& x, const Var<B>& y, const Var<B>& z) {
  return sqrt(pow(x, 2.0) + pow(y, 2.0) + pow(z, 2.0));
}

/// Dense row-major matrix of Var handles (variable_matrix.hpp). Blocks are
/// taken by copy of the handles (the nodes are shared), written with
/// set_block — the oracle's builders do not need view types.
template <class B>
class Mat {
 public:
  using V = Var<B>;
  Mat() = default;
  /// rows×cols fresh decision variables (variable_matrix.hpp:58-64).
  Mat(int rows, int cols) : m_rows{rows}, m_cols{cols} {
    m_s.reserve(rows * cols);
    for (int i = 0; i < rows * cols; ++i) m_s.emplace_back();
  }
  struct empty_t {};
  Mat(empty_t, int rows, int cols) : m_rows{rows}, m_cols{cols} {
    m_s.reserve(rows * cols);
    for (int i = 0; i < rows * cols; ++i) m_s.push_back(V::null());
  }
  Mat(std::initializer_list<std::initializer_list<V>> list) {
    m_rows = static_cast<int>(list.size());
    m_cols = m_rows ? static_cast<int>(list.begin()->size()) : 0;
    for (const auto& row : list) {
      assert(static_cast<int>(row.size()) == m_cols);
      for (const auto& v : row) m_s.push_back(v);
    }
  }
  Mat(const V& v) : m_rows{1}, m_cols{1} { m_s.push_back(v); }  // NOLINT
  /// Column vector of handles (variable_matrix.hpp:213-222).
  explicit Mat(const std::vector<V>& vars)
      : m_rows{static_cast<int>(vars.size())}, m_cols{1}, m_s{vars} {}
  /// Matrix of constants (stands in for an Eigen dense operand).
  static Mat constants(int rows, int cols, const std::vector<double>& vals) {
    Mat m{empty_t{}, rows, cols};
    for (int i = 0; i < rows * cols; ++i) m.m_s[i] = V{vals[i]};
    return m;
  }

  int rows() const { return m_rows; }
  int cols() const { return m_cols; }
  int size() const { return m_rows * m_cols; }
  V& operator()(int r, int c) { return m_s[r * m_cols + c]; }
  const V& operator()(int r, int c) const { return m_s[r * m_cols + c]; }
  V& operator[](int i) { return m_s[i]; }
  const V& operator[](int i) const { return m_s[i]; }
  auto begin() { return m_s.begin(); }
  auto end() { return m_s.end(); }
  auto begin() const { return m_s.begin(); }
  auto end() const { return m_s.end(); }

  Mat block(int r0, int c0, int nr, int nc) const {
    Mat m{empty_t{}, nr, nc};
    for (int r = 0; r < nr; ++r)
      for (int c = 0; c < nc; ++c) m(r, c) = (*this)(r0 + r, c0 + c);
    return m;
  }
  void set_block(int r0, int c0, const Mat& src) {
    for (int r = 0; r < src.rows(); ++r)
      for (int c = 0; c < src.cols(); ++c) (*this)(r0 + r, c0 + c) = src(r, c);
  }
  Mat row(int r) const { return block(r, 0, 1, m_cols); }
  Mat col(int c) const { return block(0, c, m_rows, 1); }
  Mat segment(int off, int len) const { return block(off, 0, len, 1); }
  Mat T() const {
    Mat m{empty_t{}, m_cols, m_rows};
    for (int r = 0; r < m_rows; ++r)
      for (int c = 0; c < m_cols; ++c) m(c, r) = (*this)(r, c);
    return m;
  }

  void set_value(const std::vector<double>& v) {
    for (int i = 0; i < size(); ++i) m_s[i].set_value(v[i]);
  }
  void set_value(const double* v) {
    for (int i = 0; i < size(); ++i) m_s[i].set_value(v[i]);
  }
  /// variable_matrix.hpp:993-1005: per-element Variable::value().
  std::vector<double> value() {
    std::vector<double> out(size());
    for (int i = 0; i < size(); ++i) out[i] = m_s[i].value();
    return out;
  }

  // Matrix product: left fold from constant 0 (variable_matrix.hpp:551-566).
  friend Mat operator*(const Mat& l, const Mat& r) {
    assert(l.cols() == r.rows());
    Mat out{empty_t{}, l.rows(), r.cols()};
    for (int i = 0; i < l.rows(); ++i) {
      for (int j = 0; j < r.cols(); ++j) {
        V sum{0.0};
        for (int k = 0; k < l.cols(); ++k) sum += l(i, k) * r(k, j);
        out(i, j) = sum;
      }
    }
    return out;
  }
  // scalar*matrix and matrix*scalar both build elem*scalar (:593-637).
  friend Mat operator*(const Mat& l, const V& r) {
    Mat out{empty_t{}, l.rows(), l.cols()};
    for (int i = 0; i < l.size(); ++i) out[i] = l[i] * r;
    return out;
  }
  friend Mat operator*(const V& l, const Mat& r) { return r * l; }
  friend Mat operator*(const Mat& l, double r) { return l * V{r}; }
  friend Mat operator*(double l, const Mat& r) { return r * V{l}; }
  friend Mat operator/(const Mat& l, const V& r) {  // :680-690
    Mat out{empty_t{}, l.rows(), l.cols()};
    for (int i = 0; i < l.size(); ++i) out[i] = l[i] / r;
    return out;
  }
  friend Mat operator+(const Mat& l, const Mat& r) {
    assert(l.rows() == r.rows() && l.cols() == r.cols());
    Mat out{empty_t{}, l.rows(), l.cols()};
    for (int i = 0; i < l.size(); ++i) out[i] = l[i] + r[i];
    return out;
  }
  friend Mat operator-(const Mat& l, const Mat& r) {
    assert(l.rows() == r.rows() && l.cols() == r.cols());
    Mat out{empty_t{}, l.rows(), l.cols()};
    for (int i = 0; i < l.size(); ++i) out[i] = l[i] - r[i];
    return out;
  }
  friend Mat operator-(const Mat& l) {
    Mat out{empty_t{}, l.rows(), l.cols()};
    for (int i = 0; i < l.size(); ++i) out[i] = -l[i];
    return out;
  }

 private:
  int m_rows = 0, m_cols = 0;
  std::vector<V> m_s;
};

/// Symbolic AX = B for 1x1, 2x2, 3x3 (variable_matrix.hpp:1516-1600).
template <class B>
Mat<B> solve(const Mat<B>& A, const Mat<B>& Bm) {
  assert(A.rows() == Bm.rows());
  if (A.rows() == 1 && A.cols() == 1) {
    return Mat<B>{Bm(0, 0) / A(0, 0)};
  } else if (A.rows() == 2 && A.cols() == 2) {
    const auto& a = A(0, 0);
    const auto& b = A(0, 1);
    const auto& c = A(1, 0);
    const auto& d = A(1, 1);
    Mat<B> adj{{d, -b}, {-c, a}};
    auto det = a * d - b * c;
    return adj / det * Bm;
  } else if (A.rows() == 3 && A.cols() == 3) {
    const auto& a = A(0, 0); const auto& b = A(0, 1); const auto& c = A(0, 2);
    const auto& d = A(1, 0); const auto& e = A(1, 1); const auto& f = A(1, 2);
    const auto& g = A(2, 0); const auto& h = A(2, 1); const auto& i = A(2, 2);
    auto ae = a * e; auto af = a * f; auto ah = a * h; auto ai = a * i;
    auto bd = b * d; auto bf = b * f; auto bg = b * g; auto bi = b * i;
    auto cd = c * d; auto ce = c * e; auto cg = c * g; auto ch = c * h;
    auto dh = d * h; auto di = d * i; auto eg = e * g; auto ei = e * i;
    auto fg = f * g; auto fh = f * h;
    auto adj00 = ei - fh;
    auto adj10 = fg - di;
    auto adj20 = dh - eg;
    Mat<B> adj{{adj00, ch - bi, bf - ce},
               {adj10, ai - cg, cd - af},
               {adj20, bg - ah, ae - bd}};
    auto det = a * adj00 + b * adj10 + c * adj20;
    return adj / det * Bm;
  }
  assert(false && "oracle solve(): only 1x1..3x3 are restated");
  return {};
}

/// Constraint factories: lhs − rhs per element, row-major
/// (variable.hpp:721-779); bounds(l, x, u) = {x − l, u − x} (:1006-1013).
template <class B>
std::vector<Var<B>> eq(const Mat<B>& l, const Mat<B>& r) {
  assert(l.rows() == r.rows() && l.cols() == r.cols());
  std::vector<Var<B>> out;
  for (int i = 0; i < l.size(); ++i) out.push_back(l[i] - r[i]);
  return out;
}
template <class B>
std::vector<Var<B>> eq(const Mat<B>& l, const Var<B>& r) {
  std::vector<Var<B>> out;
  for (int i = 0; i < l.size(); ++i) out.push_back(l[i] - r);
  return out;
}
template <class B>
std::vector<Var<B>> ge(const Mat<B>& l, const Mat<B>& r) { return eq(l, r); }
template <class B>
std::vector<Var<B>> ge(const Mat<B>& l, const Var<B>& r) { return eq(l, r); }
template <class B>
std::vector<Var<B>> ge(const Var<B>& l, const Mat<B>& r) {
  std::vector<Var<B>> out;
  for (int i = 0; i < r.size(); ++i) out.push_back(l - r[i]);
  return out;
}
template <class B>
std::vector<Var<B>> le(const Mat<B>& l, const Mat<B>& r) { return eq(r, l); }
template <class B>
std::vector<Var<B>> le(const Mat<B>& l, const Var<B>& r) { return ge(r, l); }
template <class B>
std::vector<Var<B>> le(const Var<B>& l, const Mat<B>& r) { return ge(r, l); }
template <class B>
std::vector<Var<B>> bounds(const Var<B>& l, const Mat<B>& x, const Var<B>& u) {
  auto out = le(l, x);
  auto hi = le(x, u);
  out.insert(out.end(), hi.begin(), hi.end());
  return out;
}

/// Symbolic reverse sweep (variable_matrix.hpp:1756-1805).
template <class B>
Mat<B> gradient_tree(const typename B::Graph& top_list, const Mat<B>& wrt) {
  using M = Mat<B>;
  assert(wrt.cols() == 1);
  if (top_list.empty()) return M{typename M::empty_t{}, wrt.rows(), 1};

  top_list[0]->adjoint_expr = B::make_const(1.0);
  for (auto* node : top_list) {
    auto& lhs = node->args[0];
    auto& rhs = node->args[1];
    if (lhs != nullptr) {
      if (rhs != nullptr) {
        lhs->adjoint_expr += node->grad_expr_l(lhs, rhs);
        rhs->adjoint_expr += node->grad_expr_r(lhs, rhs);
      } else {
        lhs->adjoint_expr += node->grad_expr_l(lhs, rhs);
      }
    }
  }
  M grad{typename M::empty_t{}, wrt.rows(), 1};
  for (int row = 0; row < grad.rows(); ++row) {
    grad[row] = Var<B>{std::move(wrt[row].expr->adjoint_expr)};
    wrt[row].expr->adjoint_expr = nullptr;
  }
  for (auto* node : top_list) node->adjoint_expr = nullptr;
  return grad;
}

}  // namespace orc
