// Solver iteration information exposed to an iteration callback (reference:
// optimization/solver/iteration_info.hpp:13-41). The references are valid only
// during the callback. The iterate lives on the device; these host mirrors are
// materialised only when at least one callback is registered.
#pragma once

#include "sleipnir/util/linalg.hpp"

namespace slp {

template <typename Scalar>
struct IterationInfo {
  /// The solver iteration.
  int iteration;
  /// The decision variables.
  const Vector<Scalar>& x;
  /// The inequality constraint slack variables.
  const Vector<Scalar>& s;
  /// The equality constraint dual variables.
  const Vector<Scalar>& y;
  /// The inequality constraint dual variables.
  const Vector<Scalar>& z;
  /// The gradient of the cost function (dense; a SparseVector in the reference).
  const Vector<Scalar>& g;
  /// The Hessian of the Lagrangian (lower triangle).
  const SparseMatrix<Scalar>& H;
  /// The equality constraint Jacobian.
  const SparseMatrix<Scalar>& A_e;
  /// The inequality constraint Jacobian.
  const SparseMatrix<Scalar>& A_i;
};

}  // namespace slp
