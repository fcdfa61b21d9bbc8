// slp::Hessian — row descriptors of ∇²variable.
//
// The Hessian is the Jacobian of the SYMBOLIC gradient (reference:
// autodiff/hessian.hpp:49-52), so construction runs detail::gradient_tree once
// and then the Jacobian row analysis (:55-103). UpLo == Lower keeps only
// row ≥ col entries when the device assembles the matrix (the reference filters
// after setFromTriplets, :151-154). value() and get() are inherited from
// Jacobian (hessian.hpp:107-157 of the reference).
#pragma once

#include <utility>

#include "sleipnir/autodiff/jacobian.hpp"

namespace slp {

/// Triangle selectors with Eigen's numeric values (Eigen::Lower = 1,
/// Eigen::Upper = 2) so existing call sites keep compiling.
enum UpLoOption : int { Lower = 1, Upper = 2 };

template <typename Scalar, int UpLo = Lower | Upper>
  requires(UpLo == Lower) || (UpLo == (Lower | Upper))
class Hessian : public Jacobian<Scalar> {
 public:
  using Base = Jacobian<Scalar>;
  Hessian(Variable<Scalar> variable, Variable<Scalar> wrt)
      : Hessian{std::move(variable), VariableMatrix<Scalar>{std::move(wrt)}} {}
  Hessian(Variable<Scalar> variable, VariableMatrix<Scalar> wrt)
      : Base{typename Base::deferred_t{},
             detail::gradient_tree(detail::topological_sort(variable.expr),
                                   wrt),
             wrt} {
    this->init();
  }
  static constexpr bool lower_only = (UpLo == Lower);

 protected:
  /// triangularView<Lower> of the evaluated matrix (hessian.hpp:151-154)
  bool keep(int row, int col) const override {
    return !lower_only || row >= col;
  }
};

}  // namespace slp
