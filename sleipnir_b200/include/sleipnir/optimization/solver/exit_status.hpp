// Solver exit status (reference: optimization/solver/exit_status.hpp:13-44;
// values are part of the API: ≥ 0 success-like, < 0 failure).
#pragma once

#include <cstdint>
#include <string_view>

namespace slp {

enum class ExitStatus : int8_t {
  SUCCESS = 0,
  CALLBACK_REQUESTED_STOP = 1,
  TOO_FEW_DOFS = -1,
  LOCALLY_INFEASIBLE = -2,
  GLOBALLY_INFEASIBLE = -3,
  FACTORIZATION_FAILED = -4,
  LINE_SEARCH_FAILED = -5,
  FEASIBILITY_RESTORATION_FAILED = -6,
  NONFINITE_INITIAL_GUESS = -7,
  DIVERGING_ITERATES = -8,
  MAX_ITERATIONS_EXCEEDED = -9,
  TIMEOUT = -10,
};

constexpr std::string_view to_string(ExitStatus s) {
  using enum ExitStatus;
  switch (s) {
    case SUCCESS: return "success";
    case CALLBACK_REQUESTED_STOP: return "callback requested stop";
    case TOO_FEW_DOFS: return "too few degrees of freedom";
    case LOCALLY_INFEASIBLE: return "locally infeasible";
    case GLOBALLY_INFEASIBLE: return "globally infeasible";
    case FACTORIZATION_FAILED: return "factorization failed";
    case LINE_SEARCH_FAILED: return "line search failed";
    case FEASIBILITY_RESTORATION_FAILED: return "feasibility restoration failed";
    case NONFINITE_INITIAL_GUESS: return "nonfinite initial guess";
    case DIVERGING_ITERATES: return "diverging iterates";
    case MAX_ITERATIONS_EXCEEDED: return "max iterations exceeded";
    case TIMEOUT: return "timeout";
  }
  return "";
}

}  // namespace slp
