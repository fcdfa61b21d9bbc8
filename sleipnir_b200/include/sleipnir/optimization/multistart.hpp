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
//
// Starts that share one KKT pattern (the usual case: the same problem from
// different guesses) CAN also share their linear algebra: the first
// Problem::solve() of every start joins a slpb_group (include/slpb.h), and the
// factorisations and triangular solves of all starts run as ONE batched launch
// per round (lane = instance, groups of 32) instead of one launch per start;
// each start's iterates are bit-identical to what it computes alone. Since the
// single-instance tree kernels became 1.8× faster in round 2 this no longer
// pays on one GPU with 16 host cores — measured on B200, cart-pole N = 5000,
// aggregate steps/s of independent launches vs grouped: 8 starts 3 628 / 2 123,
// 32 starts 3 370 / 3 083…3 608, 64 starts 4 061 / 3 667 — so it is opt-in:
// SLPB_GROUP_MIN_STARTS=k batches waves of at least k starts. (The batched
// kernels themselves, slpb_batch_*, are what many-instance callers with the
// systems already on the device use: 512 systems at 1.0 / 3.6 TB/s.)
#pragma once

#include <algorithm>
#include <cstdlib>
#include <functional>
#include <future>
#include <span>
#include <vector>

#include "sleipnir/optimization/solver/exit_status.hpp"
#include "slpb.h"

namespace slp {

/// Waves of at least this many starts batch their linear algebra by default
/// (never: see the measurements above; SLPB_GROUP_MIN_STARTS opts in).
inline constexpr int kGroupMinStarts = 1 << 30;

namespace detail {
/// What a start's thread knows about the multistart it belongs to.
struct MultistartContext {
  slpb_group* group = nullptr;
};
inline thread_local MultistartContext* tls_multistart = nullptr;
/// The start's first Problem::solve() has already answered the group's call.
inline thread_local bool tls_multistart_answered = false;
}  // namespace detail

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
        nullptr,
    int batch_device = 0) {
  using Result = MultistartResult<Scalar, DecisionVariables>;
  const size_t count = initial_guesses.size();
  std::vector<Result> results;
  results.reserve(count);
  const size_t wave =
      max_concurrency > 0 ? static_cast<size_t>(max_concurrency) : count;
  for (size_t begin = 0; begin < count; begin += wave) {
    const size_t end = std::min(count, begin + wave);
    // the starts of a wave batch their linear algebra (batch_device < 0: off)
    detail::MultistartContext context;
    const char* group_env = std::getenv("SLPB_GROUP_MIN_STARTS");
    const size_t group_min_starts =
        group_env ? static_cast<size_t>(std::max(2, std::atoi(group_env)))
                  : static_cast<size_t>(kGroupMinStarts);
    if (batch_device >= 0 && end - begin >= group_min_starts) {
      if (slpb_group_create(batch_device, static_cast<int32_t>(end - begin),
                            &context.group) != SLPB_OK) {
        context.group = nullptr;
      }
    }
    std::vector<std::future<Result>> futures;
    for (size_t i = begin; i < end; ++i) {
      futures.emplace_back(std::async(
          std::launch::async,
          [&solve, &context](const DecisionVariables& guess) -> Result {
            detail::tls_multistart = &context;
            detail::tls_multistart_answered = false;
            struct Answer {
              detail::MultistartContext& c;
              ~Answer() {
                // a start that never reached a device solve must not keep the
                // others waiting at the group's start gate
                if (c.group && !detail::tls_multistart_answered) {
                  slpb_group_abandon(c.group);
                }
                detail::tls_multistart = nullptr;
              }
            } answer{context};
            return solve(guess);
          },
          std::cref(initial_guesses[i])));
    }
    for (auto& future : futures) results.emplace_back(future.get());
    if (context.group) slpb_group_destroy(context.group);
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
