// slp::VariableMatrix / slp::VariableBlock.
//
// Same user-facing surface and graph-construction order as the reference's
// include/sleipnir/autodiff/variable_matrix.hpp and variable_block.hpp for the
// operations the direct-transcription builders use: row-major storage (:313),
// block/segment/row/col views, T() (:955-965), matrix product as a left fold
// from a constant 0 (:551-566), scalar∘matrix built as element∘scalar
// (:593-637), element-wise +/−, symbolic solve() for 1x1…3x3 (:1516-1600) and
// detail::gradient_tree (:1756-1805). Python-style Slice indexing and the
// Eigen-backed helpers (solve() for n>3, exp) are not part of the hot path and
// are not provided.
#pragma once

#include <initializer_list>
#include <span>
#include <type_traits>
#include <utility>
#include <functional>
#include <vector>

#include "sleipnir/autodiff/variable.hpp"
#include "sleipnir/util/linalg.hpp"

namespace slp {

namespace detail {
struct empty_t {};
inline constexpr empty_t empty{};
}  // namespace detail

template <typename Mat>
class VariableBlock;

/// A matrix of autodiff variables.
template <typename Scalar_>
class VariableMatrix : public SleipnirBase {
 public:
  using Scalar = Scalar_;
  using V = Variable<Scalar>;

  VariableMatrix() = default;
  /// rows x 1 fresh decision variables.
  explicit VariableMatrix(int rows) : VariableMatrix{rows, 1} {}
  /// rows x cols fresh decision variables.
  VariableMatrix(int rows, int cols) : m_rows{rows}, m_cols{cols} {
    m_storage.reserve(size_t(rows) * cols);
    for (int i = 0; i < rows * cols; ++i) m_storage.emplace_back();
  }
  /// rows x cols empty (null) variables.
  VariableMatrix(detail::empty_t, int rows, int cols)
      : m_rows{rows}, m_cols{cols} {
    m_storage.reserve(size_t(rows) * cols);
    for (int i = 0; i < rows * cols; ++i) m_storage.emplace_back(nullptr);
  }
  VariableMatrix(std::initializer_list<std::initializer_list<V>> list) {
    m_rows = static_cast<int>(list.size());
    m_cols = m_rows ? static_cast<int>(list.begin()->size()) : 0;
    for (const auto& row : list) {
      slp_assert(static_cast<int>(row.size()) == m_cols);
      m_storage.insert(m_storage.end(), row.begin(), row.end());
    }
  }
  /// Matrix of constants.
  VariableMatrix(const Matrix<Scalar>& values)  // NOLINT
      : m_rows{values.rows()}, m_cols{values.cols()} {
    for (int r = 0; r < m_rows; ++r) {
      for (int c = 0; c < m_cols; ++c) m_storage.emplace_back(values(r, c));
    }
  }
  VariableMatrix(const V& variable) : m_rows{1}, m_cols{1} {  // NOLINT
    m_storage.push_back(variable);
  }
  template <typename Mat>
  VariableMatrix(const VariableBlock<Mat>& values)  // NOLINT
      : m_rows{values.rows()}, m_cols{values.cols()} {
    for (int r = 0; r < m_rows; ++r) {
      for (int c = 0; c < m_cols; ++c) m_storage.push_back(values(r, c));
    }
  }
  /// Column vector of existing variables.
  explicit VariableMatrix(std::span<const V> values)
      : m_rows{static_cast<int>(values.size())}, m_cols{1},
        m_storage(values.begin(), values.end()) {}
  VariableMatrix(std::span<const V> values, int rows, int cols)
      : m_rows{rows}, m_cols{cols}, m_storage(values.begin(), values.end()) {
    slp_assert(static_cast<int>(values.size()) == rows * cols);
  }

  int rows() const { return m_rows; }
  int cols() const { return m_cols; }
  int size() const { return m_rows * m_cols; }

  V& operator()(int r, int c) {
    slp_assert(r >= 0 && r < m_rows && c >= 0 && c < m_cols);
    return m_storage[size_t(r) * m_cols + c];
  }
  const V& operator()(int r, int c) const {
    slp_assert(r >= 0 && r < m_rows && c >= 0 && c < m_cols);
    return m_storage[size_t(r) * m_cols + c];
  }
  V& operator[](int r, int c) { return (*this)(r, c); }
  const V& operator[](int r, int c) const { return (*this)(r, c); }
  V& operator[](int i) { return m_storage[i]; }
  const V& operator[](int i) const { return m_storage[i]; }
  V& operator()(int i) { return m_storage[i]; }
  const V& operator()(int i) const { return m_storage[i]; }

  VariableBlock<VariableMatrix> block(int r0, int c0, int nr, int nc) {
    return VariableBlock<VariableMatrix>{*this, r0, c0, nr, nc};
  }
  VariableBlock<const VariableMatrix> block(int r0, int c0, int nr,
                                            int nc) const {
    return VariableBlock<const VariableMatrix>{*this, r0, c0, nr, nc};
  }
  auto segment(int offset, int length) {
    slp_assert(m_cols == 1);
    return block(offset, 0, length, 1);
  }
  auto segment(int offset, int length) const {
    slp_assert(m_cols == 1);
    return block(offset, 0, length, 1);
  }
  auto row(int r) { return block(r, 0, 1, m_cols); }
  auto row(int r) const { return block(r, 0, 1, m_cols); }
  auto col(int c) { return block(0, c, m_rows, 1); }
  auto col(int c) const { return block(0, c, m_rows, 1); }

  VariableMatrix T() const {
    VariableMatrix out{detail::empty, m_cols, m_rows};
    for (int r = 0; r < m_rows; ++r) {
      for (int c = 0; c < m_cols; ++c) out(c, r) = (*this)(r, c);
    }
    return out;
  }

  /// Applies a unary operator to every element, row-major
  /// (variable_matrix.hpp:1027-1039).
  VariableMatrix cwise_transform(
      const std::function<V(const V& x)>& unary_op) const {
    VariableMatrix out{detail::empty, m_rows, m_cols};
    for (int r = 0; r < m_rows; ++r) {
      for (int c = 0; c < m_cols; ++c) out(r, c) = unary_op((*this)(r, c));
    }
    return out;
  }

  void set_value(const Matrix<Scalar>& values) {
    slp_assert(values.rows() == m_rows && values.cols() == m_cols);
    for (int r = 0; r < m_rows; ++r) {
      for (int c = 0; c < m_cols; ++c) (*this)(r, c).set_value(values(r, c));
    }
  }
  /// Column-vector overload used by the solver to push x back into the leaves.
  void set_value(const Scalar* values) {
    for (size_t i = 0; i < m_storage.size(); ++i) {
      m_storage[i].set_value(values[i]);
    }
  }
  Scalar value(int r, int c) { return (*this)(r, c).value(); }
  Scalar value(int i) { return m_storage[i].value(); }
  Matrix<Scalar> value() {
    Matrix<Scalar> out{m_rows, m_cols};
    for (int r = 0; r < m_rows; ++r) {
      for (int c = 0; c < m_cols; ++c) out(r, c) = value(r, c);
    }
    return out;
  }

  auto begin() { return m_storage.begin(); }
  auto end() { return m_storage.end(); }
  auto begin() const { return m_storage.begin(); }
  auto end() const { return m_storage.end(); }

 private:
  int m_rows = 0, m_cols = 0;
  std::vector<V> m_storage;
};

/// A rectangular view into a VariableMatrix.
template <typename Mat>
class VariableBlock : public SleipnirBase {
 public:
  using Scalar = typename std::remove_const_t<Mat>::Scalar;
  using V = Variable<Scalar>;

  VariableBlock(Mat& mat, int r0, int c0, int nr, int nc)
      : m_mat{&mat}, m_r0{r0}, m_c0{c0}, m_rows{nr}, m_cols{nc} {
    slp_assert(r0 >= 0 && c0 >= 0 && nr >= 0 && nc >= 0);
    slp_assert(r0 + nr <= mat.rows() && c0 + nc <= mat.cols());
  }
  VariableBlock(const VariableBlock&) = default;

  /// Assigning a block writes through to the viewed matrix.
  VariableBlock& operator=(const VariableBlock& values) {
    if (this != &values) assign(values);
    return *this;
  }
  template <MatrixLike M>
  VariableBlock& operator=(const M& values) {
    assign(values);
    return *this;
  }
  VariableBlock& operator=(ScalarLike auto value) {
    slp_assert(m_rows == 1 && m_cols == 1);
    (*this)(0, 0) = value;
    return *this;
  }

  int rows() const { return m_rows; }
  int cols() const { return m_cols; }
  int size() const { return m_rows * m_cols; }

  decltype(auto) operator()(int r, int c) const {
    slp_assert(r >= 0 && r < m_rows && c >= 0 && c < m_cols);
    return (*m_mat)(m_r0 + r, m_c0 + c);
  }
  decltype(auto) operator[](int r, int c) const { return (*this)(r, c); }
  decltype(auto) operator()(int i) const {
    return (*this)(i / m_cols, i % m_cols);
  }
  decltype(auto) operator[](int i) const { return (*this)(i); }

  VariableBlock block(int r0, int c0, int nr, int nc) const {
    return VariableBlock{*m_mat, m_r0 + r0, m_c0 + c0, nr, nc};
  }
  VariableBlock segment(int offset, int length) const {
    slp_assert(m_cols == 1);
    return block(offset, 0, length, 1);
  }
  VariableBlock row(int r) const { return block(r, 0, 1, m_cols); }
  VariableBlock col(int c) const { return block(0, c, m_rows, 1); }

  VariableMatrix<Scalar> T() const { return VariableMatrix<Scalar>{*this}.T(); }
  VariableMatrix<Scalar> cwise_transform(
      const std::function<V(const V& x)>& unary_op) const {
    return VariableMatrix<Scalar>{*this}.cwise_transform(unary_op);
  }

  void set_value(const Matrix<Scalar>& values) const {
    for (int r = 0; r < m_rows; ++r) {
      for (int c = 0; c < m_cols; ++c) (*this)(r, c).set_value(values(r, c));
    }
  }
  void set_value(Scalar value) const {
    slp_assert(m_rows == 1 && m_cols == 1);
    (*this)(0, 0).set_value(value);
  }
  Scalar value(int r, int c) const {
    return const_cast<V&>((*this)(r, c)).value();
  }
  Scalar value(int i) const { return value(i / m_cols, i % m_cols); }
  Matrix<Scalar> value() const {
    Matrix<Scalar> out{m_rows, m_cols};
    for (int r = 0; r < m_rows; ++r) {
      for (int c = 0; c < m_cols; ++c) out(r, c) = value(r, c);
    }
    return out;
  }

 private:
  template <typename M>
  void assign(const M& values) {
    slp_assert(values.rows() == m_rows && values.cols() == m_cols);
    for (int r = 0; r < m_rows; ++r) {
      for (int c = 0; c < m_cols; ++c) (*this)(r, c) = V{values(r, c)};
    }
  }

  Mat* m_mat;
  int m_r0, m_c0, m_rows, m_cols;
};

// ---- arithmetic --------------------------------------------------------------
// Every overload below builds nodes in the same order as the reference so that
// the resulting graphs (and hence rounding) coincide.

namespace detail {
template <typename T>
concept AnyMatrix = MatrixLike<T>;
template <typename L, typename R>
concept MatrixPair =
    AnyMatrix<L> && AnyMatrix<R> && (SleipnirType<L> || SleipnirType<R>);
template <typename L, typename R>
using pair_scalar_t = constraint_scalar_t<L, R>;
}  // namespace detail

/// Matrix product: each entry is a left fold `sum += l(i,k) * r(k,j)` that
/// starts from a constant 0 (which the `+` pruning rule then drops).
template <typename L, typename R>
  requires detail::MatrixPair<L, R>
auto operator*(const L& lhs, const R& rhs) {
  using Scalar = detail::pair_scalar_t<L, R>;
  slp_assert(lhs.cols() == rhs.rows());
  VariableMatrix<Scalar> out{detail::empty, lhs.rows(), rhs.cols()};
  for (int i = 0; i < lhs.rows(); ++i) {
    for (int j = 0; j < rhs.cols(); ++j) {
      Variable<Scalar> sum{Scalar(0)};
      for (int k = 0; k < lhs.cols(); ++k) {
        sum += Variable<Scalar>{lhs(i, k)} * Variable<Scalar>{rhs(k, j)};
      }
      out(i, j) = sum;
    }
  }
  return out;
}

/// matrix ∘ scalar and scalar ∘ matrix both build `element * scalar`.
template <SleipnirMatrixLike L, ScalarLike R>
auto operator*(const L& lhs, const R& rhs) {
  using Scalar = typename L::Scalar;
  VariableMatrix<Scalar> out{detail::empty, lhs.rows(), lhs.cols()};
  const Variable<Scalar> s{rhs};
  for (int r = 0; r < lhs.rows(); ++r) {
    for (int c = 0; c < lhs.cols(); ++c) out(r, c) = lhs(r, c) * s;
  }
  return out;
}
template <ScalarLike L, SleipnirMatrixLike R>
auto operator*(const L& lhs, const R& rhs) {
  return rhs * lhs;
}
template <NumericMatrixLike L, typename Scalar>
auto operator*(const L& lhs, const Variable<Scalar>& rhs) {
  VariableMatrix<Scalar> out{detail::empty, lhs.rows(), lhs.cols()};
  for (int r = 0; r < lhs.rows(); ++r) {
    for (int c = 0; c < lhs.cols(); ++c) {
      out(r, c) = Variable<Scalar>{lhs(r, c)} * rhs;
    }
  }
  return out;
}
template <typename Scalar, NumericMatrixLike R>
auto operator*(const Variable<Scalar>& lhs, const R& rhs) {
  return rhs * lhs;
}

template <SleipnirMatrixLike L, ScalarLike R>
auto operator/(const L& lhs, const R& rhs) {
  using Scalar = typename L::Scalar;
  VariableMatrix<Scalar> out{detail::empty, lhs.rows(), lhs.cols()};
  const Variable<Scalar> s{rhs};
  for (int r = 0; r < lhs.rows(); ++r) {
    for (int c = 0; c < lhs.cols(); ++c) out(r, c) = lhs(r, c) / s;
  }
  return out;
}
template <NumericMatrixLike L, typename Scalar>
auto operator/(const L& lhs, const Variable<Scalar>& rhs) {
  VariableMatrix<Scalar> out{detail::empty, lhs.rows(), lhs.cols()};
  for (int r = 0; r < lhs.rows(); ++r) {
    for (int c = 0; c < lhs.cols(); ++c) {
      out(r, c) = Variable<Scalar>{lhs(r, c)} / rhs;
    }
  }
  return out;
}

template <typename L, typename R>
  requires detail::MatrixPair<L, R>
auto operator+(const L& lhs, const R& rhs) {
  using Scalar = detail::pair_scalar_t<L, R>;
  slp_assert(lhs.rows() == rhs.rows() && lhs.cols() == rhs.cols());
  VariableMatrix<Scalar> out{detail::empty, lhs.rows(), lhs.cols()};
  for (int r = 0; r < lhs.rows(); ++r) {
    for (int c = 0; c < lhs.cols(); ++c) {
      out(r, c) = Variable<Scalar>{lhs(r, c)} + Variable<Scalar>{rhs(r, c)};
    }
  }
  return out;
}
template <typename L, typename R>
  requires detail::MatrixPair<L, R>
auto operator-(const L& lhs, const R& rhs) {
  using Scalar = detail::pair_scalar_t<L, R>;
  slp_assert(lhs.rows() == rhs.rows() && lhs.cols() == rhs.cols());
  VariableMatrix<Scalar> out{detail::empty, lhs.rows(), lhs.cols()};
  for (int r = 0; r < lhs.rows(); ++r) {
    for (int c = 0; c < lhs.cols(); ++c) {
      out(r, c) = Variable<Scalar>{lhs(r, c)} - Variable<Scalar>{rhs(r, c)};
    }
  }
  return out;
}
template <SleipnirMatrixLike M>
auto operator-(const M& m) {
  using Scalar = typename M::Scalar;
  VariableMatrix<Scalar> out{detail::empty, m.rows(), m.cols()};
  for (int r = 0; r < m.rows(); ++r) {
    for (int c = 0; c < m.cols(); ++c) out(r, c) = -m(r, c);
  }
  return out;
}

template <typename Scalar, MatrixLike R>
VariableMatrix<Scalar>& operator+=(VariableMatrix<Scalar>& lhs, const R& rhs) {
  slp_assert(lhs.rows() == rhs.rows() && lhs.cols() == rhs.cols());
  for (int r = 0; r < lhs.rows(); ++r) {
    for (int c = 0; c < lhs.cols(); ++c) {
      lhs(r, c) += Variable<Scalar>{rhs(r, c)};
    }
  }
  return lhs;
}
template <typename Scalar, MatrixLike R>
VariableMatrix<Scalar>& operator-=(VariableMatrix<Scalar>& lhs, const R& rhs) {
  slp_assert(lhs.rows() == rhs.rows() && lhs.cols() == rhs.cols());
  for (int r = 0; r < lhs.rows(); ++r) {
    for (int c = 0; c < lhs.cols(); ++c) {
      lhs(r, c) -= Variable<Scalar>{rhs(r, c)};
    }
  }
  return lhs;
}

/// Solves A X = B symbolically for 1x1, 2x2 and 3x3 A via the adjugate.
template <typename Scalar>
VariableMatrix<Scalar> solve(const VariableMatrix<Scalar>& A,
                             const VariableMatrix<Scalar>& B) {
  using VM = VariableMatrix<Scalar>;
  slp_assert(A.rows() == B.rows());
  if (A.rows() == 1 && A.cols() == 1) {
    return VM{B(0, 0) / A(0, 0)};
  } else if (A.rows() == 2 && A.cols() == 2) {
    const auto& a = A(0, 0);
    const auto& b = A(0, 1);
    const auto& c = A(1, 0);
    const auto& d = A(1, 1);
    VM adj_A{{d, -b}, {-c, a}};
    auto det_A = a * d - b * c;
    return adj_A / det_A * B;
  } else if (A.rows() == 3 && A.cols() == 3) {
    const auto& a = A(0, 0); const auto& b = A(0, 1); const auto& c = A(0, 2);
    const auto& d = A(1, 0); const auto& e = A(1, 1); const auto& f = A(1, 2);
    const auto& g = A(2, 0); const auto& h = A(2, 1); const auto& i = A(2, 2);
    auto ae = a * e; auto af = a * f; auto ah = a * h; auto ai = a * i;
    auto bd = b * d; auto bf = b * f; auto bg = b * g; auto bi = b * i;
    auto cd = c * d; auto ce = c * e; auto cg = c * g; auto ch = c * h;
    auto dh = d * h; auto di = d * i; auto eg = e * g; auto ei = e * i;
    auto fg = f * g; auto fh = f * h;
    auto adj_A00 = ei - fh;
    auto adj_A10 = fg - di;
    auto adj_A20 = dh - eg;
    VM adj_A{{adj_A00, ch - bi, bf - ce},
             {adj_A10, ai - cg, cd - af},
             {adj_A20, bg - ah, ae - bd}};
    auto det_A = a * adj_A00 + b * adj_A10 + c * adj_A20;
    return adj_A / det_A * B;
  }
  slp_assert(false && "solve(): only 1x1, 2x2 and 3x3 systems are supported");
  return VM{};
}

namespace detail {

/// Symbolic reverse sweep over `top_list` (parent→child): returns ∂root/∂wrt
/// as expressions; an element is empty when wrt[i] is not in the graph.
template <typename Scalar>
VariableMatrix<Scalar> gradient_tree(const ExpressionGraph& top_list,
                                     const VariableMatrix<Scalar>& wrt) {
  slp_assert(wrt.cols() == 1);
  VariableMatrix<Scalar> grad{detail::empty, wrt.rows(), 1};
  if (top_list.empty()) return grad;

  // Adjoint expressions, keyed by node id (nodes created during the sweep get
  // ids beyond the initial pool size and never need one).
  std::vector<Expr> adjoint_expr(pool().size());
  adjoint_expr[top_list[0]] = constant_ptr(1.0);
  for (ExprId node : top_list) {
    // Re-read through pool(): the sweep appends nodes and may reallocate.
    const ExprId l = pool().lhs[node], r = pool().rhs[node];
    if (l == kNull) continue;
    const Op op = static_cast<Op>(pool().op[node]);
    const Expr& ae = adjoint_expr[node];
    const Expr le{l};
    if (r != kNull) {
      const Expr re{r};
      adjoint_expr[l] = adjoint_expr[l] + grad_expr_l(op, ae, le, re);
      adjoint_expr[r] = adjoint_expr[r] + grad_expr_r(op, ae, le, re);
    } else {
      adjoint_expr[l] = adjoint_expr[l] + grad_expr_l(op, ae, le, Expr{});
    }
  }
  for (int row = 0; row < grad.rows(); ++row) {
    const ExprId id = wrt(row).expr.id();
    if (id < static_cast<ExprId>(adjoint_expr.size())) {
      grad(row) = Variable<Scalar>{std::move(adjoint_expr[id])};
    }
  }
  return grad;
}

}  // namespace detail

}  // namespace slp
