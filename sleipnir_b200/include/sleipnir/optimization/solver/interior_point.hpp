// Host driver of the device-resident interior-point method.
//
// Control flow, constants and exit conditions follow the reference's
// include/sleipnir/optimization/solver/interior_point.hpp:123-866 line by line
// (citations inline); every vector/matrix operation of that loop is a call
// into the C ABI (include/slpb.h) and runs on the GPU. Per call only a handful
// of scalars return to the host, where the branchy logic lives: the
// inertia-correcting δ/γ loop (solver/util/sparse_regularized_ldlt.hpp:64-152),
// the filter (solver/util/filter.hpp:70-212), second-order corrections, the
// barrier-parameter schedule and the exit statuses.
#pragma once

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cmath>
#include <functional>
#include <limits>
#include <span>
#include <stdexcept>
#include <string>
#include <vector>

#include "sleipnir/optimization/solver/exit_status.hpp"
#include "sleipnir/optimization/solver/iteration_info.hpp"
#include "sleipnir/optimization/solver/options.hpp"
#include "sleipnir/util/print_diagnostics.hpp"
#include "sleipnir/util/linalg.hpp"
#include "slpb.h"

namespace slp {

/// Thrown when the device library itself fails (no GPU, CUDA error, bad
/// upload). Numerical outcomes are reported through ExitStatus, never thrown.
class DeviceError : public std::runtime_error {
 public:
  using std::runtime_error::runtime_error;
};

namespace detail {
inline void device_check(int rc, const slpb_solver* s, const char* what) {
  if (rc != SLPB_OK) {
    throw DeviceError(std::string{what} + ": " +
                      (s ? slpb_last_error(s) : "no device handle") +
                      " (status " + std::to_string(rc) + ")");
  }
}
}  // namespace detail
#define SLP_DEVICE_CALL(handle, call) \
  ::slp::detail::device_check((call), (handle), #call)

/// One row per Newton iteration (diagnostics and parity tests); mirrors the
/// columns of the reference's iteration table
/// (util/print_diagnostics.hpp:193-237).
struct IterationRecord {
  int iteration = 0;
  int type = 0;  ///< 0 normal, 1 feasibility restoration
  double error = 0, cost = 0, infeasibility = 0, complementarity = 0;
  double mu = 0, delta = 0, gamma = 0, alpha = 0, alpha_max = 0, alpha_z = 0;
  int factorizations = 0, solves = 0, trials = 0;
  double t_end = 0.0;  ///< seconds since the solve entered the solver
  std::vector<double> x, s, y, z;  ///< filled when SolveTrace::keep_iterates
};

struct SolveTrace {
  bool keep_iterates = false;
  /// Benchmark hygiene: evict the device L2 (slpb_flush_l2) at the start of
  /// every iteration; the time spent flushing is excluded from t_end and
  /// loop_seconds.
  bool flush_l2 = false;
  double flush_seconds = 0.0;
  std::vector<IterationRecord> rows;
  int64_t factorizations = 0, solves = 0, trials = 0;
  double loop_seconds = 0.0;  ///< wall time inside the Newton loop
};

/// FilterEntry / Filter: solver/util/filter.hpp:20-212.
template <typename Scalar>
struct FilterEntry {
  Scalar cost{0};
  Scalar constraint_violation{0};
  constexpr FilterEntry() = default;
  constexpr FilterEntry(Scalar cost, Scalar constraint_violation)
      : cost{cost}, constraint_violation{constraint_violation} {}
  constexpr bool dominated_by(const FilterEntry& e) const {
    return e.cost <= cost && e.constraint_violation <= constraint_violation;
  }
};

template <typename Scalar>
class Filter {
 public:
  Scalar min_constraint_violation;
  Scalar max_constraint_violation;

  explicit Filter(Scalar initial_constraint_violation = Scalar(0)) {
    min_constraint_violation =
        Scalar(1e-4) * std::max(Scalar(1), initial_constraint_violation);
    max_constraint_violation =
        Scalar(1e4) * std::max(Scalar(1), initial_constraint_violation);
  }
  void reset() {
    m_filter.clear();
    m_last_rejection_due_to_filter = false;
  }
  bool try_add(const FilterEntry<Scalar>& current,
               const FilterEntry<Scalar>& trial, Scalar D_phi, Scalar alpha) {
    using std::isfinite;
    using std::pow;
    if (!isfinite(trial.cost) ||
        trial.constraint_violation > max_constraint_violation) {
      return false;
    }
    constexpr Scalar s_phi(2.3), s_theta(1.1), eta_phi(1e-8);
    const bool switching =
        D_phi < Scalar(0) &&
        alpha * pow(-D_phi, s_phi) > pow(current.constraint_violation, s_theta);
    const bool armijo = trial.cost <= current.cost + eta_phi * alpha * D_phi;
    const Scalar phi = pow(alpha, Scalar(1.5));
    const bool sufficient_decrease =
        trial.cost <=
            current.cost - phi * kGammaCost * current.constraint_violation ||
        trial.constraint_violation <=
            (Scalar(1) - phi * kGammaConstraint) *
                current.constraint_violation;
    if (current.constraint_violation <= min_constraint_violation &&
        switching) {
      if (!armijo) {
        m_last_rejection_due_to_filter = false;
        return false;
      }
    } else if (!sufficient_decrease) {
      m_last_rejection_due_to_filter = false;
      return false;
    }
    if (std::any_of(m_filter.begin(), m_filter.end(),
                    [&](const auto& e) { return trial.dominated_by(e); })) {
      m_last_rejection_due_to_filter = true;
      return false;
    }
    if (!switching || !armijo) {
      const FilterEntry<Scalar> entry{
          current.cost - phi * kGammaCost * current.constraint_violation,
          (Scalar(1) - phi * kGammaConstraint) * current.constraint_violation};
      std::erase_if(m_filter,
                    [&](const auto& e) { return e.dominated_by(entry); });
      m_filter.push_back(entry);
    }
    return true;
  }
  bool last_rejection_due_to_filter() const {
    return m_last_rejection_due_to_filter;
  }

 private:
  static constexpr Scalar kGammaCost{1e-8};
  static constexpr Scalar kGammaConstraint{1e-5};
  std::vector<FilterEntry<Scalar>> m_filter;
  bool m_last_rejection_due_to_filter = false;
};

/// The inertia-correcting regularisation loop around the device factorisation
/// (solver/util/sparse_regularized_ldlt.hpp:64-152). The matrix stays resident
/// on the device: a retry only changes δ and γ.
class DeviceRegularizedLDLT {
 public:
  DeviceRegularizedLDLT(slpb_solver* dev, int num_decision_variables,
                        int num_equality_constraints, double gamma_min)
      : m_dev{dev}, m_n{num_decision_variables},
        m_me{num_equality_constraints}, m_gamma_min{gamma_min} {}

  /// Returns false on NumericalIssue (δ or γ beyond 1e20).
  ///
  /// The decision sequence is the reference's: try (0, 0); then (δ, γ_min) with
  /// δ from the previous solve; then grow δ/γ by the inertia seen. The second
  /// candidate is known before the first attempt returns, so when the previous
  /// iteration needed regularisation both are factored side by side in one
  /// launch (slpb_factor_pair) and the one the sequential loop would have kept
  /// is selected. `factorizations` counts the attempts of the sequential
  /// algorithm (what the reference would have run).
  bool compute() {
    slpb_factor_info info{};
    double delta = m_prev_delta == 0.0
                       ? 1e-4
                       : std::max(m_prev_delta / 2.0,
                                  std::numeric_limits<double>::epsilon());
    double gamma = m_gamma_min;
    bool have_second = false;
    slpb_factor_info second{};
    if (m_speculate && m_prev_delta != 0.0) {
      const double d2[2] = {0.0, delta}, g2[2] = {0.0, gamma};
      slpb_factor_info pair[2]{};
      ++factorizations;
      SLP_DEVICE_CALL(m_dev, slpb_factor_pair(m_dev, d2, g2, 1, pair));
      info = pair[0];
      second = pair[1];
      have_second = true;
    } else {
      factor(0.0, 0.0, /*reassemble=*/1, info);
    }
    if (!info.zero_pivot && ideal(info) && info.min_abs_d >= 1e-4) {
      m_prev_delta = 0.0;
      m_prev_gamma = 0.0;
      return true;
    }
    while (true) {
      if (have_second) {
        have_second = false;
        ++factorizations;
        info = second;
        SLP_DEVICE_CALL(m_dev, slpb_select_factor(m_dev, 1));
      } else {
        factor(delta, gamma, /*reassemble=*/0, info);
      }
      if (!info.zero_pivot) {
        if (ideal(info)) {
          m_prev_delta = delta;
          m_prev_gamma = gamma;
          return true;
        } else if (info.n_zero > 0) {
          if (gamma == 0.0) {
            gamma = 1e-10;
          } else {
            delta *= 10.0;
            gamma *= 10.0;
          }
        } else if (info.n_neg > m_me) {
          delta *= 10.0;
        } else if (info.n_pos > m_n) {
          gamma = gamma == 0.0 ? 1e-10 : gamma * 10.0;
        }
      } else {
        delta *= 10.0;
        gamma = gamma == 0.0 ? 1e-10 : gamma * 10.0;
      }
      if (delta > 1e20 || gamma > 1e20) {
        m_prev_delta = delta;
        m_prev_gamma = gamma;
        return false;
      }
    }
  }

  double hessian_regularization() const { return m_prev_delta; }
  double constraint_jacobian_regularization() const { return m_prev_gamma; }
  /// Test hook: seeds "previous δ" so a single step can be replayed.
  void set_previous_regularization(double delta, double gamma) {
    m_prev_delta = delta;
    m_prev_gamma = gamma;
  }
  int factorizations = 0;
  /// Enables the side-by-side factorisation of the first two candidates.
  void set_speculation(bool on) { m_speculate = on; }

 private:
  bool ideal(const slpb_factor_info& i) const {
    return i.n_pos == m_n && i.n_neg == m_me && i.n_zero == 0;
  }
  void factor(double delta, double gamma, int reassemble,
              slpb_factor_info& info) {
    ++factorizations;
    SLP_DEVICE_CALL(m_dev, slpb_factor(m_dev, delta, gamma, reassemble, &info));
  }

  slpb_solver* m_dev;
  int m_n, m_me;
  double m_gamma_min;
  double m_prev_delta = 0.0, m_prev_gamma = 0.0;
  bool m_speculate = true;
};

namespace detail {

enum class KKTErrorType { INF_NORM_SCALED, ONE_NORM };

/// kkt_error.hpp:92-146 from the reductions the device returned.
inline double kkt_error_scaled(const slpb_kkt_stats& k, int me, int mi,
                               double mu) {
  constexpr double s_max = 100.0;
  const double s_d =
      std::max(s_max, (k.y_l1 + k.z_l1) / double(me + mi)) / s_max;
  const double s_c = std::max(s_max, k.z_l1 / double(mi)) / s_max;
  // the device reports ±∞ extrema for an empty s∘z
  const double sz_inf =
      mi == 0 ? 0.0
              : std::max(std::abs(k.sz_max - mu), std::abs(k.sz_min - mu));
  return std::max({k.r_inf / s_d, sz_inf / s_c, k.ce_inf, k.cis_inf});
}
/// kkt_error.hpp:216-251 (μ = 0, as the solver uses it for E_0).
inline double kkt_error_unscaled_mu0(const slpb_kkt_stats& k, int me, int mi) {
  constexpr double s_max = 100.0;
  const double s_d =
      std::max(s_max, (k.u_y_l1 + k.u_z_l1) / double(me + mi)) / s_max;
  const double s_c = std::max(s_max, k.u_z_l1 / double(mi)) / s_max;
  const double sz_inf =
      mi == 0 ? 0.0 : std::max(std::abs(k.u_sz_max), std::abs(k.u_sz_min));
  return std::max({k.u_r_inf / s_d, sz_inf / s_c, k.u_ce_inf, k.u_cis_inf});
}
inline double kkt_error_one_norm(const slpb_kkt_stats& k) {
  return k.r_l1 + k.sz_mu_l1 + k.ce_l1 + k.cis_l1;
}

}  // namespace detail

/// The reference has three Newton-type drivers of one shape: interior_point.hpp,
/// sqp.hpp (no slacks, no barrier, one step length for x and y, restoration
/// entered with μ = tolerance/10) and newton.hpp (no constraints, no
/// second-order correction, no restoration, factorisation failures ignored).
/// Problem::solve picks one by the constraint counts (problem.hpp:335, :403,
/// :512); here they are one loop over the same device kernels.
enum class SolverKind { IPM, SQP, NEWTON };

/// Dimensions and scaling the driver needs besides the device handle.
struct DeviceProblemInfo {
  int num_decision_variables = 0;
  int num_equality_constraints = 0;
  int num_inequality_constraints = 0;
  double scaling_f = 1.0;
  /// Ranks solving this problem in lockstep (slpb_comm_init); > 1 makes the
  /// rank-local stop decisions collective (slpb_comm_agree).
  int world = 1;
};

/// Hook used by Problem::solve to run feasibility restoration on a second
/// device problem; returns the restoration's ExitStatus
/// (feasibility_restoration.hpp:346-628).
/// accept_test(trial_x, trial_s): the stopping rule of the restoration phase
/// (interior_point.hpp:738-756) evaluated on the ORIGINAL problem — the trial
/// reduces the violation to < 0.9× its value at entry and the outer filter
/// accepts it.
using RestorationAcceptTest = std::function<bool(
    const std::vector<double>& trial_x, const std::vector<double>& trial_s)>;
using RestorationHook = std::function<ExitStatus(
    double mu, int& iterations, const RestorationAcceptTest& accept_test)>;

/// Finds the optimal solution to a nonlinear program with the interior-point
/// method, with the iterate resident on the device (reference overload 2,
/// interior_point.hpp:123-134).
template <typename Scalar>
ExitStatus interior_point(
    slpb_solver* dev, const DeviceProblemInfo& problem,
    std::span<std::function<bool(const IterationInfo<Scalar>& info)>>
        iteration_callbacks,
    const Options& options, bool in_feasibility_restoration, Scalar& mu,
    int& iterations, SolveTrace* trace = nullptr,
    const RestorationHook* restoration = nullptr,
    double initial_delta = 0.0, SolverKind kind = SolverKind::IPM) {
  using std::isfinite;
  const auto solve_start_time = std::chrono::steady_clock::now();
  const int n = problem.num_decision_variables;
  const int me = problem.num_equality_constraints;
  const int mi = problem.num_inequality_constraints;

  // f, g, H, c_e, A_e, c_i, A_i at the initial iterate (:245-251)
  slpb_point_info cur{};
  SLP_DEVICE_CALL(dev, slpb_eval_current(dev, 1, &cur));

  if (me > n) return ExitStatus::TOO_FEW_DOFS;  // :274-280

  constexpr int kAllFinite = SLPB_FINITE_F | SLPB_FINITE_C_E |
                             SLPB_FINITE_C_I | SLPB_FINITE_G |
                             SLPB_FINITE_A_E | SLPB_FINITE_A_I | SLPB_FINITE_H;
  if ((cur.finite & kAllFinite) != kAllFinite) {
    return ExitStatus::NONFINITE_INITIAL_GUESS;  // :283-286
  }

  const Scalar mu_min = problem.scaling_f * Scalar(options.tolerance) / 10.0;
  constexpr Scalar tau_min(0.99);
  Scalar tau = tau_min;

  Filter<Scalar> filter{cur.ce_l1 + cur.cis_l1};  // :303-304

  auto update_barrier_parameter_and_reset_filter = [&] {  // :308-333
    constexpr Scalar kappa_mu(0.2);
    constexpr Scalar theta_mu(1.5);
    mu = std::max(mu_min, std::min(kappa_mu * mu, std::pow(mu, theta_mu)));
    tau = std::max(tau_min, Scalar(1) - mu);
    filter.reset();
  };

  // The reference picks Eigen's dense LDLT when the lower triangle is ≥ 25 %
  // full (:340-348); the device path always factors sparse.
  // γ_min = 0 inside feasibility restoration (:352). (The device ordering keeps
  // every multiplier behind one of its neighbours — symbolic.cpp,
  // defer_leading_multipliers — so γ = 0 does not meet structurally zero
  // pivots on the relaxed problem.)
  DeviceRegularizedLDLT solver{dev, n, me,
                               in_feasibility_restoration ? 0.0 : 1e-10};
  solver.set_previous_regularization(initial_delta, 0.0);
  if (std::getenv("SLPB_NO_SPECULATION")) solver.set_speculation(false);
  // slpb_solve_trial / slpb_accept_relinearize: three host round trips per
  // iteration instead of five (development switch: the call-per-step path)
  const bool merged_calls = std::getenv("SLPB_NO_MERGED_CALLS") == nullptr;

  constexpr Scalar alpha_reduction_factor(0.5);
  constexpr Scalar alpha_min(1e-7);
  int full_step_rejected_counter = 0;

  slpb_kkt_stats kkt{};
  SLP_DEVICE_CALL(dev, slpb_kkt_stats_current(dev, 0.0, &kkt));
  Scalar E_0 = detail::kkt_error_unscaled_mu0(kkt, me, mi);  // :361-362

  // Host mirrors for IterationInfo, filled lazily.
  Vector<Scalar> hx, hs, hy, hz, hg;
  SparseMatrix<Scalar> hH, hAe, hAi;
  auto download_matrix = [&](int which_pattern, int which_values) {
    int32_t rows = 0, cols = 0;
    int64_t nnz = 0;
    SLP_DEVICE_CALL(dev, slpb_pattern(dev, which_pattern, &rows, &cols, &nnz,
                                      nullptr, nullptr));
    std::vector<int32_t> outer(cols + 1), inner(nnz);
    std::vector<Scalar> values(nnz);
    SLP_DEVICE_CALL(dev, slpb_pattern(dev, which_pattern, &rows, &cols, &nnz,
                                      outer.data(), inner.data()));
    if (nnz > 0) {
      SLP_DEVICE_CALL(dev, slpb_download(dev, which_values, values.data()));
    }
    return SparseMatrix<Scalar>{rows, cols, std::move(outer), std::move(inner),
                                std::move(values)};
  };
  auto download_iterate = [&] {
    hx.resize(n);
    hs.resize(mi);
    hy.resize(me);
    hz.resize(mi);
    SLP_DEVICE_CALL(dev, slpb_get_iterate(dev, hx.data(), hs.data(), hy.data(),
                                          hz.data()));
  };

  // (development switch, read once per solve — not inside the Newton loop)
  const bool fused_forward = std::getenv("SLPB_NO_FUSED_FORWARD") == nullptr;
  const auto loop_start_time = std::chrono::steady_clock::now();
  struct LoopTimer {
    SolveTrace* t;
    std::chrono::steady_clock::time_point t0;
    ~LoopTimer() {
      if (t) {
        t->loop_seconds += std::chrono::duration<double>(
                               std::chrono::steady_clock::now() - t0)
                               .count() -
                           t->flush_seconds;
      }
    }
    // the restoration sub-solve runs inside the outer loop's time
  } loop_timer{in_feasibility_restoration ? nullptr : trace, loop_start_time};

  while (E_0 > Scalar(options.tolerance)) {
    if (trace && trace->flush_l2) {
      const auto f0 = std::chrono::steady_clock::now();
      SLP_DEVICE_CALL(dev, slpb_flush_l2(dev));
      trace->flush_seconds += std::chrono::duration<double>(
                                  std::chrono::steady_clock::now() - f0)
                                  .count();
    }
    int it_solves = 0, it_trials = 0;
    const int fact_before = solver.factorizations;
    const auto iteration_start_time = std::chrono::steady_clock::now();

    // Local infeasibility (:387-402; is_locally_infeasible.hpp:17-60)
    if (me > 0 && kkt.aetce_l2 < 1e-6 && kkt.ce_l2 > 1e-2) {
      return ExitStatus::LOCALLY_INFEASIBLE;
    }
    if (mi > 0 && kkt.aitcip_l2 < 1e-6 && kkt.cip_l2 > 1e-6) {
      return ExitStatus::LOCALLY_INFEASIBLE;
    }
    // Diverging iterates (:405-408)
    if (kkt.x_inf > 1e10 || kkt.s_inf > 1e10 || !kkt.xs_finite) {
      return ExitStatus::DIVERGING_ITERATES;
    }

    // Iteration callbacks (:414-418)
    if (!iteration_callbacks.empty()) {
      download_iterate();
      hg.resize(n);
      SLP_DEVICE_CALL(dev, slpb_download(dev, SLPB_ARR_G, hg.data()));
      hH = download_matrix(SLPB_OUT_H_C, SLPB_ARR_H_VAL);
      hAe = download_matrix(SLPB_OUT_A_E, SLPB_ARR_A_E_VAL);
      hAi = download_matrix(SLPB_OUT_A_I, SLPB_ARR_A_I_VAL);
      bool stop = false;
      for (const auto& callback : iteration_callbacks) {
        if (callback({iterations, hx, hs, hy, hz, hg, hH, hAe, hAi})) {
          stop = true;
          break;
        }
      }
      if (problem.world > 1) {
        int32_t any = 0;
        SLP_DEVICE_CALL(dev, slpb_comm_agree(dev, stop ? 1 : 0, &any));
        stop = any != 0;
      }
      if (stop) return ExitStatus::CALLBACK_REQUESTED_STOP;
    }

    Scalar alpha_max(1), alpha(1), alpha_z(1);
    bool call_feasibility_restoration = false;
    bool relinearized = false;

    // lhs assembly + factorisation with inertia correction (:426-465)
    // The right-hand side does not depend on δ/γ: build it first so that the
    // factorisation carries its forward substitution (slpb_prepare_rhs).
    if (fused_forward) {
      SLP_DEVICE_CALL(dev, slpb_prepare_rhs(dev, mu));
    }
    if (!solver.compute() && kind != SolverKind::NEWTON) {
      return ExitStatus::FACTORIZATION_FAILED;  // newton.hpp:183 carries on
    }

    // rhs, solve, step recovery, fraction-to-the-boundary (:444-497) — and,
    // in the same host round trip, the trial point at the full step
    // (slpb_solve_trial): the first trial of the line search below.
    slpb_step_info step{};
    slpb_point_info trial{};
    const int slack_from_ci =
        options.feasible_ipm && cur.ci_all_positive ? 1 : 0;
    bool have_full_step_trial = false;
    if (merged_calls) {
      SLP_DEVICE_CALL(dev, slpb_solve_trial(dev, mu, tau,
                                            kind != SolverKind::IPM ? 1 : 0,
                                            slack_from_ci, &step, &trial));
      have_full_step_trial = true;
    } else {
      SLP_DEVICE_CALL(dev, slpb_solve(dev, mu, tau, &step));
    }
    ++it_solves;
    alpha_max = step.alpha_max;
    alpha = alpha_max;
    if (alpha < alpha_min) call_feasibility_restoration = true;
    // sqp.hpp:346 moves y by the primal step length
    alpha_z = kind == SolverKind::IPM ? step.alpha_z : alpha;

    const FilterEntry<Scalar> current_entry{cur.f - mu * cur.log_s_sum,
                                            cur.ce_l1 + cur.cis_l1};  // :499
    const Scalar D_phi = step.g_dot_px - mu * step.sinv_dot_ps;       // :508

    slpb_step_info accepted_step = step;
    while (true) {  // :512-717
      ++it_trials;
      if (kind != SolverKind::IPM) alpha_z = alpha;
      if (have_full_step_trial) {
        have_full_step_trial = false;  // α = α_max: already evaluated
      } else {
        SLP_DEVICE_CALL(
            dev, slpb_trial(dev, alpha, alpha_z, 0, slack_from_ci, &trial));
      }

      constexpr int kTrialFinite =
          SLPB_FINITE_F | SLPB_FINITE_C_E | SLPB_FINITE_C_I;
      if ((trial.finite & kTrialFinite) != kTrialFinite) {  // :532-542
        alpha *= alpha_reduction_factor;
        if (alpha < alpha_min) {
          if (kind == SolverKind::NEWTON) {
            return ExitStatus::LINE_SEARCH_FAILED;  // newton.hpp:213
          }
          call_feasibility_restoration = true;
          break;
        }
        continue;
      }

      const FilterEntry<Scalar> trial_entry{trial.f - mu * trial.log_s_sum,
                                            trial.ce_l1 + trial.cis_l1};
      if (filter.try_add(current_entry, trial_entry, D_phi, alpha)) break;

      const Scalar prev_constraint_violation = cur.ce_l1 + cur.cis_l1;
      Scalar next_constraint_violation = trial.ce_l1 + trial.cis_l1;

      // Second-order corrections (:561-664)
      if (kind != SolverKind::NEWTON && alpha == alpha_max &&
          next_constraint_violation >= prev_constraint_violation) {
        Scalar alpha_soc = alpha;
        Scalar alpha_z_soc = alpha_z;
        Scalar soc_constraint_violation = next_constraint_violation;
        bool step_acceptable = false;
        SLP_DEVICE_CALL(dev, slpb_soc_begin(dev));
        for (int soc_iteration = 0; soc_iteration < 5 && !step_acceptable;
             ++soc_iteration) {
          slpb_step_info soc_step{};
          SLP_DEVICE_CALL(
              dev, slpb_soc_iterate(dev, mu, tau, alpha_soc, &soc_step));
          ++it_solves;
          alpha_soc = soc_step.alpha_max;
          alpha_z_soc =
              kind == SolverKind::IPM ? soc_step.alpha_z : alpha_soc;
          ++it_trials;
          SLP_DEVICE_CALL(
              dev, slpb_trial(dev, alpha_soc, alpha_z_soc, 1, 0, &trial));
          const FilterEntry<Scalar> soc_entry{trial.f - mu * trial.log_s_sum,
                                              trial.ce_l1 + trial.cis_l1};
          // NB: the acceptance test uses the OUTER α (:637)
          if (filter.try_add(current_entry, soc_entry, D_phi, alpha)) {
            accepted_step = soc_step;
            alpha = alpha_soc;
            alpha_z = alpha_z_soc;
            step_acceptable = true;
            break;
          }
          constexpr Scalar kappa_soc(0.99);
          next_constraint_violation = trial.ce_l1 + trial.cis_l1;
          if (next_constraint_violation >
              kappa_soc * soc_constraint_violation) {
            break;
          }
          soc_constraint_violation = next_constraint_violation;
        }
        if (step_acceptable) break;
      }

      if (alpha == alpha_max) ++full_step_rejected_counter;  // :669-671

      // Filter reset heuristic (:677-684)
      if (kind != SolverKind::NEWTON && full_step_rejected_counter >= 4 &&
          filter.max_constraint_violation >
              current_entry.constraint_violation / Scalar(10) &&
          filter.last_rejection_due_to_filter()) {
        filter.max_constraint_violation *= Scalar(0.1);
        filter.reset();
        continue;
      }

      alpha *= alpha_reduction_factor;

      // Step size hit the minimum: accept anyway if the KKT error went down,
      // else restore feasibility (:691-716)
      if (alpha < alpha_min) {
        slpb_kkt_stats ks{};
        SLP_DEVICE_CALL(dev, slpb_kkt_stats_current(dev, mu, &ks));
        const Scalar current_kkt_error = detail::kkt_error_one_norm(ks);
        ++it_trials;
        if (kind != SolverKind::IPM) alpha_z = alpha_max;
        SLP_DEVICE_CALL(dev, slpb_trial(dev, alpha_max, alpha_z, 0, 0, &trial));
        SLP_DEVICE_CALL(dev, slpb_kkt_stats_trial(dev, mu, &ks));
        const Scalar next_kkt_error = detail::kkt_error_one_norm(ks);
        if (next_kkt_error <= Scalar(0.999) * current_kkt_error) break;
        if (kind == SolverKind::NEWTON) {
          return ExitStatus::LINE_SEARCH_FAILED;  // newton.hpp:245
        }
        call_feasibility_restoration = true;
        break;
      }
    }

    if (call_feasibility_restoration) {  // :721-771
      if (in_feasibility_restoration || restoration == nullptr) {
        return ExitStatus::FEASIBILITY_RESTORATION_FAILED;
      }
      // The trial-point KKT probe above may have overwritten the derivative
      // arrays; bring them back to the current iterate for the hook.
      SLP_DEVICE_CALL(dev, slpb_eval_current(dev, 2, &cur));
      const FilterEntry<Scalar> initial_entry = current_entry;
      // outer iterate and gradient at entry, for D_ϕ of the acceptance test
      std::vector<double> x0(n), s0(mi), g0(n);
      SLP_DEVICE_CALL(dev, slpb_get_iterate(dev, x0.data(), s0.data(), nullptr,
                                            nullptr));
      SLP_DEVICE_CALL(dev, slpb_download(dev, SLPB_ARR_G, g0.data()));
      const RestorationAcceptTest accept_test =
          [&](const std::vector<double>& trial_x,
              const std::vector<double>& trial_s) -> bool {
        slpb_point_info probe{};
        SLP_DEVICE_CALL(dev, slpb_probe_point(dev, trial_x.data(),
                                              trial_s.data(), &probe));
        const FilterEntry<Scalar> trial_entry{
            probe.f - mu * probe.log_s_sum, probe.ce_l1 + probe.cis_l1};
        Scalar gdx(0), sds(0);
        for (int i = 0; i < n; ++i) gdx += g0[i] * (trial_x[i] - x0[i]);
        for (int i = 0; i < mi; ++i) {
          sds += (Scalar(1) / s0[i]) * (trial_s[i] - s0[i]);
        }
        const Scalar D_phi_restoration = gdx - mu * sds;
        return trial_entry.constraint_violation <
                   Scalar(0.9) * initial_entry.constraint_violation &&
               filter.try_add(initial_entry, trial_entry, D_phi_restoration,
                              alpha);
      };
      // the SQP variant enters with μ = tolerance/10
      // (feasibility_restoration.hpp:121)
      // restoration runs on another handle for many iterations: the other
      // members of a batching group must not wait for this one meanwhile
      SLP_DEVICE_CALL(dev, slpb_group_pause(dev));
      struct Resume {
        slpb_solver* d;
        ~Resume() { slpb_group_resume(d); }
      } resume{dev};
      ExitStatus status = (*restoration)(
          kind == SolverKind::IPM ? mu : Scalar(options.tolerance) / 10.0,
          iterations, accept_test);
      if (status != ExitStatus::SUCCESS) return status;
      SLP_DEVICE_CALL(dev, slpb_eval_current(dev, 0, &cur));
    } else {
      if (alpha == alpha_max) full_step_rejected_counter = 0;  // :774-776
      // x, s, y, z, f, c_e, c_i ← trial; clamp z (:779-805)
      if (merged_calls) {
        // … re-linearise and reduce the errors in the same round trip
        int32_t finite = 0;
        SLP_DEVICE_CALL(dev, slpb_accept_relinearize(dev, mu, &finite, &kkt));
        cur = trial;
        cur.finite = finite;
        relinearized = true;
      } else {
        SLP_DEVICE_CALL(dev, slpb_accept(dev, mu));
        cur = trial;
      }
    }

    if (!relinearized) {
      // Re-linearise: A_e, A_i, g, H at the new iterate (:809-812)
      slpb_point_info derivs{};
      SLP_DEVICE_CALL(dev, slpb_eval_current(dev, 2, &derivs));
      cur.finite = derivs.finite;
      // error reductions (:815-832)
      SLP_DEVICE_CALL(dev, slpb_kkt_stats_current(dev, mu, &kkt));
    }

    // E_0 and the barrier update (:815-832)
    E_0 = detail::kkt_error_unscaled_mu0(kkt, me, mi);
    if (kind == SolverKind::IPM && E_0 > Scalar(options.tolerance)) {
      constexpr Scalar kappa_eps(10);
      Scalar E_mu = detail::kkt_error_scaled(kkt, me, mi, mu);
      while (mu > mu_min && E_mu <= kappa_eps * mu) {
        update_barrier_parameter_and_reset_filter();
        E_mu = detail::kkt_error_scaled(kkt, me, mi, mu);
      }
    }

    if (options.diagnostics) {
      // the table of print_diagnostics.hpp:193-239 (interior_point.hpp:837-849);
      // sᵀz = ‖Sz − 0·e‖₁ since every sᵢzᵢ > 0. (Second-order-correction
      // sub-rows are not printed: their error column would need the
      // derivatives at the rejected trial point.)
      slpb_kkt_stats at_zero{};
      SLP_DEVICE_CALL(dev, slpb_kkt_stats_current(dev, 0.0, &at_zero));
      print_iteration_diagnostics(
          iterations,
          in_feasibility_restoration ? IterationType::FEASIBILITY_RESTORATION
                                     : IterationType::NORMAL,
          std::chrono::duration<double, std::milli>(
              std::chrono::steady_clock::now() - iteration_start_time)
              .count(),
          E_0, cur.f, cur.ce_l1 + cur.cis_l1, mi > 0 ? at_zero.sz_mu_l1 : 0.0,
          mu, solver.hessian_regularization(),
          solver.constraint_jacobian_regularization(),
          std::max(accepted_step.px_inf, accepted_step.ps_inf),
          std::max(accepted_step.py_inf, accepted_step.pz_inf), alpha,
          alpha_max, alpha_reduction_factor, alpha_z);
    }

    if (trace != nullptr) {
      IterationRecord row;
      row.iteration = iterations;
      row.type = in_feasibility_restoration ? 1 : 0;
      row.error = E_0;
      row.cost = cur.f;
      row.infeasibility = cur.ce_l1 + cur.cis_l1;
      row.complementarity = 0.0;
      row.mu = mu;
      row.delta = solver.hessian_regularization();
      row.gamma = solver.constraint_jacobian_regularization();
      row.alpha = alpha;
      row.alpha_max = alpha_max;
      row.alpha_z = alpha_z;
      row.factorizations = solver.factorizations - fact_before;
      row.solves = it_solves;
      row.trials = it_trials;
      row.t_end = std::chrono::duration<double>(
                      std::chrono::steady_clock::now() - solve_start_time)
                      .count() -
                  trace->flush_seconds;
      if (trace->keep_iterates) {
        row.x.resize(n);
        row.s.resize(mi);
        row.y.resize(me);
        row.z.resize(mi);
        SLP_DEVICE_CALL(dev, slpb_get_iterate(dev, row.x.data(), row.s.data(),
                                              row.y.data(), row.z.data()));
      }
      trace->factorizations += row.factorizations;
      trace->solves += it_solves;
      trace->trials += it_trials;
      trace->rows.push_back(std::move(row));
    }
    (void)accepted_step;

    ++iterations;
    if (iterations >= options.max_iterations) {  // :855-857
      return ExitStatus::MAX_ITERATIONS_EXCEEDED;
    }
    bool timed_out = std::chrono::steady_clock::now() - solve_start_time >
                     options.timeout;  // :860-862
    if (problem.world > 1) {
      // every rank reads its own clock: leave together or not at all
      int32_t any = 0;
      SLP_DEVICE_CALL(dev, slpb_comm_agree(dev, timed_out ? 1 : 0, &any));
      timed_out = any != 0;
    }
    if (timed_out) return ExitStatus::TIMEOUT;
  }
  return ExitStatus::SUCCESS;
}

}  // namespace slp
