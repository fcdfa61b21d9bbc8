// Solver options (reference: optimization/solver/options.hpp:13-38).
#pragma once

#include <chrono>
#include <limits>

namespace slp {

struct Options {
  /// The solver will stop once the error is below this tolerance.
  double tolerance = 1e-8;
  /// The maximum number of solver iterations before returning a solution.
  int max_iterations = 5000;
  /// The maximum elapsed wall clock time before returning a solution.
  std::chrono::duration<double> timeout{
      std::numeric_limits<double>::infinity()};
  /// Enables the feasible interior-point method (slacks follow c_i once all
  /// inequality constraints are feasible).
  bool feasible_ipm = false;
  /// Enables diagnostic output.
  bool diagnostics = false;
};

}  // namespace slp
