// slp::multistart — the reference's multi-instance driver
// (include/sleipnir/optimization/multistart.hpp:16-73), same surface: the user
// hands in a function that builds and solves one problem from one initial
// guess, and a list of guesses; the result with the lowest cost among the
// successful solves wins.
//
// On the device path every solve owns a handle (its own CUDA stream and
// device arrays) and the expression pool is thread-local, so the starts are
// independent end to end. That is what makes this the data-parallel axis of
// the library: one Newton step is latency-bound (a dependent chain through the
// assembly tree) and leaves most of the GPU idle, so kernels of
// concurrent starts overlap on the SMs. `max_concurrency` bounds how many
// run at once (0: all of them, as the reference does).
#pragma once

#include <algorithm>
#include <functional>
#include <future>
#include <span>
#include <vector>

#include "sleipnir/optimization/solver/exit_status.hpp"

namespace slp {

/// The result of a multistart solve (multistart.hpp:16-29).
template <typename Scalar, typename DecisionVariables>
struct MultistartResult {
  /// The solver exit status.
  ExitStatus status;
  /// The solution's cost.
  Scalar cost;
  /// The decision variables.
  DecisionVariables variables;
};

/// Solves an optimization problem from different starting points in parallel,
/// each on its own thread, and returns the solution with the lowest cost;
/// successful solves are preferred over unsuccessful ones
/// (multistart.hpp:31-73).
template <typename Scalar, typename DecisionVariables>
MultistartResult<Scalar, DecisionVariables> multistart(
    const std::function<MultistartResult<Scalar, DecisionVariables>(
        const DecisionVariables& initial_guess)>& solve,
    std::span<const DecisionVariables> initial_guesses,
    int max_concurrency = 0,
    std::vector<MultistartResult<Scalar, DecisionVariables>>* all_results =
        nullptr) {
  using Result = MultistartResult<Scalar, DecisionVariables>;
  const size_t count = initial_guesses.size();
  std::vector<Result> results;
  results.reserve(count);
  const size_t wave =
      max_concurrency > 0 ? static_cast<size_t>(max_concurrency) : count;
  for (size_t begin = 0; begin < count; begin += wave) {
    std::vector<std::future<Result>> futures;
    for (size_t i = begin; i < std::min(count, begin + wave); ++i) {
      futures.emplace_back(std::async(std::launch::async, std::cref(solve),
                                      std::cref(initial_guesses[i])));
    }
    for (auto& future : futures) results.emplace_back(future.get());
  }
  if (all_results) *all_results = results;

  return *std::ranges::min_element(
      results, [](const Result& a, const Result& b) {
        // Prioritize successful solve
        if (a.status == ExitStatus::SUCCESS && b.status != ExitStatus::SUCCESS) {
          return true;
        }
        // Otherwise prioritize solution with lower cost
        return a.cost < b.cost;
      });
}

}  // namespace slp
