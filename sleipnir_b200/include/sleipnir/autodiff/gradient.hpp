// slp::Gradient — a one-row Jacobian (reference: autodiff/gradient.hpp:53-57).
#pragma once

#include <utility>

#include "sleipnir/autodiff/jacobian.hpp"

namespace slp {

template <typename Scalar>
class Gradient {
 public:
  Gradient(Variable<Scalar> variable, Variable<Scalar> wrt)
      : m_jacobian{std::move(variable), std::move(wrt)} {}
  Gradient(Variable<Scalar> variable, VariableMatrix<Scalar> wrt)
      : m_jacobian{std::move(variable), std::move(wrt)} {}
  const Jacobian<Scalar>& jacobian() const { return m_jacobian; }

  /// The gradient as a column of expressions (gradient.hpp:63-65).
  VariableMatrix<Scalar> get() const { return m_jacobian.get().T(); }

  /// The gradient at wrt's current values (gradient.hpp:72-80; the reference
  /// returns an Eigen::SparseVector, this a dense column).
  const Vector<Scalar>& value() {
    const SparseMatrix<Scalar>& J = m_jacobian.value();
    m_g = Vector<Scalar>(J.cols());
    for (int c = 0; c < J.cols(); ++c) m_g[c] = J.coeff(0, c);
    return m_g;
  }

 private:
  Jacobian<Scalar> m_jacobian;
  Vector<Scalar> m_g;
};

}  // namespace slp
