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

 private:
  Jacobian<Scalar> m_jacobian;
};

}  // namespace slp
