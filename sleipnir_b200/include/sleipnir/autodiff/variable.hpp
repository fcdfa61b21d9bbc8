// slp::Variable and the constraint factories.
//
// Same user-facing surface and graph-construction semantics as the reference's
// include/sleipnir/autodiff/variable.hpp (default-constructed Variable is a
// fresh decision variable :271-272; set_value/value :125-151; arithmetic and
// math overloads; `==`, `>=`, `<=`, bounds() turn into lhs − rhs expressions
// :721-779, :896-1013), built on the flat expression pool.
#pragma once

#include <algorithm>
#include <concepts>
#include <initializer_list>
#include <utility>
#include <vector>

#include "sleipnir/autodiff/expression.hpp"
#include "sleipnir/autodiff/expression_graph.hpp"
#include "sleipnir/autodiff/expression_type.hpp"
#include "sleipnir/autodiff/sleipnir_base.hpp"
#include "sleipnir/util/assert.hpp"
#include "sleipnir/util/concepts.hpp"

namespace slp {

template <typename Scalar>
class VariableMatrix;

/// An autodiff variable pointing to an expression node.
template <typename Scalar_>
class Variable : public SleipnirBase {
 public:
  using Scalar = Scalar_;
  static_assert(std::same_as<Scalar, double>,
                "the device path computes in FP64 only");

  /// Constructs a linear Variable with a value of zero (a decision variable).
  Variable() : expr{detail::decision_variable_ptr()} {}
  /// Constructs an empty Variable.
  explicit Variable(std::nullptr_t) : expr{nullptr} {}
  /// Constant from a floating-point or integral value.
  Variable(std::floating_point auto value)  // NOLINT
      : expr{detail::constant_ptr(static_cast<double>(value))} {}
  Variable(std::integral auto value)  // NOLINT
      : expr{detail::constant_ptr(static_cast<double>(value))} {}
  /// From a 1x1 matrix-like DSL object.
  template <SleipnirMatrixLike M>
  Variable(const M& value) : expr{value(0, 0).expr} {  // NOLINT
    slp_assert(value.rows() == 1 && value.cols() == 1);
  }
  explicit Variable(const detail::Expr& e) : expr{e} {}
  explicit Variable(detail::Expr&& e) : expr{std::move(e)} {}

  /// Assignment of a number makes this Variable a constant.
  Variable& operator=(ScalarLike auto value)
    requires(!SleipnirType<decltype(value)>)
  {
    expr = detail::constant_ptr(static_cast<double>(value));
    m_graph_initialized = false;
    return *this;
  }

  /// Sets the value of a decision variable.
  void set_value(Scalar value) { expr.set_val(value); }

  /// Returns the value of this variable, re-evaluating its expression graph on
  /// the host (user-facing query; the solver's own sweeps run on the device).
  Scalar value() {
    if (!m_graph_initialized) {
      m_graph = detail::topological_sort(expr);
      m_graph_initialized = true;
    }
    detail::update_values(m_graph);
    return expr.val();
  }

  ExpressionType type() const { return expr.type(); }

  friend Variable operator*(const Variable& l, const Variable& r) {
    return Variable{l.expr * r.expr};
  }
  friend Variable operator/(const Variable& l, const Variable& r) {
    return Variable{l.expr / r.expr};
  }
  friend Variable operator+(const Variable& l, const Variable& r) {
    return Variable{l.expr + r.expr};
  }
  friend Variable operator-(const Variable& l, const Variable& r) {
    return Variable{l.expr - r.expr};
  }
  friend Variable operator-(const Variable& l) { return Variable{-l.expr}; }
  friend Variable operator+(const Variable& l) { return Variable{+l.expr}; }
  Variable& operator*=(const Variable& r) { return *this = *this * r; }
  Variable& operator/=(const Variable& r) { return *this = *this / r; }
  Variable& operator+=(const Variable& r) { return *this = *this + r; }
  Variable& operator-=(const Variable& r) { return *this = *this - r; }

  /// The expression node.
  detail::Expr expr;

 private:
  detail::ExpressionGraph m_graph;
  bool m_graph_initialized = false;
};

template <std::floating_point T>
Variable(T) -> Variable<T>;
template <std::integral T>
Variable(T) -> Variable<double>;

#define SLP_UNARY_FN(fn)                              \
  template <typename Scalar>                          \
  Variable<Scalar> fn(const Variable<Scalar>& x) {    \
    return Variable<Scalar>{detail::fn(x.expr)};      \
  }
SLP_UNARY_FN(abs)
SLP_UNARY_FN(acos)
SLP_UNARY_FN(asin)
SLP_UNARY_FN(atan)
SLP_UNARY_FN(cbrt)
SLP_UNARY_FN(cos)
SLP_UNARY_FN(cosh)
SLP_UNARY_FN(erf)
SLP_UNARY_FN(exp)
SLP_UNARY_FN(log)
SLP_UNARY_FN(log10)
SLP_UNARY_FN(sign)
SLP_UNARY_FN(sin)
SLP_UNARY_FN(sinh)
SLP_UNARY_FN(sqrt)
SLP_UNARY_FN(tan)
SLP_UNARY_FN(tanh)
#undef SLP_UNARY_FN

#define SLP_BINARY_FN(fn)                                                     \
  template <typename Scalar>                                                  \
  Variable<Scalar> fn(const Variable<Scalar>& a, const Variable<Scalar>& b) { \
    return Variable<Scalar>{detail::fn(a.expr, b.expr)};                      \
  }                                                                           \
  template <typename Scalar>                                                  \
  Variable<Scalar> fn(const Variable<Scalar>& a, const ScalarLike auto& b)    \
    requires(!SleipnirType<decltype(b)>)                                      \
  {                                                                           \
    return Variable<Scalar>{detail::fn(a.expr, Variable<Scalar>(b).expr)};    \
  }                                                                           \
  template <typename Scalar>                                                  \
  Variable<Scalar> fn(const ScalarLike auto& a, const Variable<Scalar>& b)    \
    requires(!SleipnirType<decltype(a)>)                                      \
  {                                                                           \
    return Variable<Scalar>{detail::fn(Variable<Scalar>(a).expr, b.expr)};    \
  }
SLP_BINARY_FN(atan2)
SLP_BINARY_FN(hypot)
SLP_BINARY_FN(max)
SLP_BINARY_FN(min)
SLP_BINARY_FN(pow)
#undef SLP_BINARY_FN

/// hypot(x, y, z) = sqrt(x² + y² + z²).
template <typename Scalar>
Variable<Scalar> hypot(const Variable<Scalar>& x, const Variable<Scalar>& y,
                       const Variable<Scalar>& z) {
  return sqrt(pow(x, 2) + pow(y, 2) + pow(z, 2));
}

namespace detail {

template <typename T>
struct scalar_of {
  using type = double;
};
template <SleipnirType T>
struct scalar_of<T> {
  using type = typename std::decay_t<T>::Scalar;
};

/// Element access that works for DSL matrices, numeric matrices and scalars.
template <typename T>
decltype(auto) elem(const T& v, int r, int c) {
  if constexpr (MatrixLike<T>) {
    return v(r, c);
  } else {
    return v;
  }
}

}  // namespace detail

/// lhs − rhs per element, row-major; the standard form is c(x) = 0 / c(x) ≥ 0.
template <typename Scalar, typename LHS, typename RHS>
std::vector<Variable<Scalar>> make_constraints(const LHS& lhs, const RHS& rhs) {
  std::vector<Variable<Scalar>> constraints;
  if constexpr (MatrixLike<LHS> && MatrixLike<RHS>) {
    slp_assert(lhs.rows() == rhs.rows() && lhs.cols() == rhs.cols());
  }
  int rows = 1, cols = 1;
  if constexpr (MatrixLike<LHS>) {
    rows = lhs.rows();
    cols = lhs.cols();
  } else if constexpr (MatrixLike<RHS>) {
    rows = rhs.rows();
    cols = rhs.cols();
  }
  constraints.reserve(size_t(rows) * cols);
  for (int r = 0; r < rows; ++r) {
    for (int c = 0; c < cols; ++c) {
      constraints.emplace_back(Variable<Scalar>{detail::elem(lhs, r, c)} -
                               Variable<Scalar>{detail::elem(rhs, r, c)});
    }
  }
  return constraints;
}

/// A vector of equality constraints of the form c(x) = 0.
template <typename Scalar>
struct EqualityConstraints {
  std::vector<Variable<Scalar>> constraints;

  EqualityConstraints(std::initializer_list<EqualityConstraints> list) {
    for (const auto& e : list) {
      constraints.insert(constraints.end(), e.constraints.begin(),
                         e.constraints.end());
    }
  }
  explicit EqualityConstraints(const std::vector<EqualityConstraints>& list) {
    for (const auto& e : list) {
      constraints.insert(constraints.end(), e.constraints.begin(),
                         e.constraints.end());
    }
  }
  template <typename LHS, typename RHS>
    requires(SleipnirType<LHS> || SleipnirType<RHS>)
  EqualityConstraints(const LHS& lhs, const RHS& rhs)
      : constraints{make_constraints<Scalar>(lhs, rhs)} {}

  /// True if every constraint is satisfied at the current values.
  operator bool() {  // NOLINT
    return std::ranges::all_of(constraints,
                               [](auto& c) { return c.value() == Scalar(0); });
  }
};

/// A vector of inequality constraints of the form c(x) ≥ 0.
template <typename Scalar>
struct InequalityConstraints {
  std::vector<Variable<Scalar>> constraints;

  InequalityConstraints(std::initializer_list<InequalityConstraints> list) {
    for (const auto& e : list) {
      constraints.insert(constraints.end(), e.constraints.begin(),
                         e.constraints.end());
    }
  }
  explicit InequalityConstraints(
      const std::vector<InequalityConstraints>& list) {
    for (const auto& e : list) {
      constraints.insert(constraints.end(), e.constraints.begin(),
                         e.constraints.end());
    }
  }
  template <typename LHS, typename RHS>
    requires(SleipnirType<LHS> || SleipnirType<RHS>)
  InequalityConstraints(const LHS& lhs, const RHS& rhs)
      : constraints{make_constraints<Scalar>(lhs, rhs)} {}

  operator bool() {  // NOLINT
    return std::ranges::all_of(constraints,
                               [](auto& c) { return c.value() >= Scalar(0); });
  }
};

namespace detail {
template <typename LHS, typename RHS>
using constraint_scalar_t =
    std::conditional_t<SleipnirType<LHS>, typename scalar_of<LHS>::type,
                       typename scalar_of<RHS>::type>;
template <typename T>
concept ConstraintOperand = ScalarLike<T> || MatrixLike<T>;
}  // namespace detail

template <detail::ConstraintOperand LHS, detail::ConstraintOperand RHS>
  requires(SleipnirType<LHS> || SleipnirType<RHS>)
auto operator==(const LHS& lhs, const RHS& rhs) {
  return EqualityConstraints<detail::constraint_scalar_t<LHS, RHS>>{lhs, rhs};
}
template <detail::ConstraintOperand LHS, detail::ConstraintOperand RHS>
  requires(SleipnirType<LHS> || SleipnirType<RHS>)
auto operator>=(const LHS& lhs, const RHS& rhs) {
  return InequalityConstraints<detail::constraint_scalar_t<LHS, RHS>>{lhs, rhs};
}
template <detail::ConstraintOperand LHS, detail::ConstraintOperand RHS>
  requires(SleipnirType<LHS> || SleipnirType<RHS>)
auto operator<=(const LHS& lhs, const RHS& rhs) {
  return rhs >= lhs;
}
template <detail::ConstraintOperand LHS, detail::ConstraintOperand RHS>
  requires(SleipnirType<LHS> || SleipnirType<RHS>)
auto operator<(const LHS& lhs, const RHS& rhs) {
  return rhs >= lhs;
}
template <detail::ConstraintOperand LHS, detail::ConstraintOperand RHS>
  requires(SleipnirType<LHS> || SleipnirType<RHS>)
auto operator>(const LHS& lhs, const RHS& rhs) {
  return lhs >= rhs;
}

/// l ≤ x ≤ u as {x − l ≥ 0, u − x ≥ 0}.
template <detail::ConstraintOperand L, SleipnirType X,
          detail::ConstraintOperand U>
auto bounds(const L& l, const X& x, const U& u) {
  using Scalar = typename detail::scalar_of<X>::type;
  return InequalityConstraints<Scalar>{l <= x, x <= u};
}

}  // namespace slp
