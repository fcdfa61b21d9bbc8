// slp::Problem — the user-facing DSL and solver entry point.
//
// Same surface as the reference's include/sleipnir/optimization/problem.hpp:
// decision_variable (:78-101), symmetric_decision_variable (:118-140),
// minimize / maximize (:151-190), subject_to (:196-234), solve (:281),
// add_callback ×2 / clear_callbacks / add_persistent_callback (:690-730).
// solve() builds the same autodiff objects as the reference's IPM branch
// (:517-560), detects conflicting bounds (:597-606), computes the problem
// scaling (:615-616), then hands the flattened graphs to the device library
// (include/slpb.h) and runs the interior-point loop there. The reference
// dispatches problems without inequality constraints to its Newton / SQP
// solvers (:335,403); those are outside this build, so every problem takes
// the interior-point path (an empty inequality set is allowed).
#pragma once

#include <algorithm>
#include <concepts>
#include <array>
#include <chrono>
#include <stdexcept>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <limits>
#include <memory>
#include <optional>
#include <span>
#include <utility>
#include <vector>

#include "sleipnir/autodiff/gradient.hpp"
#include "sleipnir/autodiff/hessian.hpp"
#include "sleipnir/autodiff/jacobian.hpp"
#include "sleipnir/autodiff/variable.hpp"
#include "sleipnir/autodiff/variable_matrix.hpp"
#include "sleipnir/optimization/solver/device_problem.hpp"
#include "sleipnir/optimization/solver/exit_status.hpp"
#include "sleipnir/optimization/multistart.hpp"
#include "sleipnir/optimization/solver/interior_point.hpp"
#include "sleipnir/optimization/solver/iteration_info.hpp"
#include "sleipnir/optimization/solver/options.hpp"
#include "sleipnir/util/print_diagnostics.hpp"
#include "sleipnir/util/spy.hpp"
#include "slpb.h"

namespace slp {

/// Device-side knobs that have no counterpart in the reference's Options.
struct DeviceOptions {
  int device = 0;                               ///< CUDA ordinal
  int ordering = SLPB_ORDER_NESTED_DISSECTION;  ///< slpb_ordering
  std::vector<int32_t> permutation;             ///< for SLPB_ORDER_CUSTOM
  bool keep_iterates = false;                   ///< record x,s,y,z per iteration
  bool flush_l2 = false;  ///< evict L2 before every iteration (benchmarks)
  /// slpb_factor_arithmetic of the LDLᵀ (−1: the library's default, i.e.
  /// SLPB_ARITH_REFERENCE unless SLPB_FACTOR_ARITH=tensor is set).
  int factor_arithmetic = -1;
  /// Write H.spy, A_e.spy, A_i.spy (sparsity pattern + signs per iteration,
  /// util/spy.hpp) into the working directory, as the reference's
  /// solve(options, spy = true) does (problem.hpp:365-380, :462-480, :562-595).
  bool spy = false;
  /// Multi-GPU: one process per GPU, all solving the SAME problem in lockstep;
  /// the re-linearisation sweep is sharded and exchanged with one NCCL
  /// all-gather per iteration (slpb_comm_init). world == 1: single GPU.
  int rank = 0, world = 1;
  std::array<char, 128> nccl_unique_id{};
};

/// RAII owner of a device handle.
struct DeviceHandle {
  slpb_solver* s = nullptr;
  explicit DeviceHandle(int device) {
    const int rc = slpb_create(device, &s);
    if (rc != SLPB_OK) {
      throw DeviceError(
          "slpb_create failed (status " + std::to_string(rc) +
          "): the interior-point path needs a CUDA device; there is no CPU "
          "fallback");
    }
  }
  DeviceHandle(const DeviceHandle&) = delete;
  DeviceHandle& operator=(const DeviceHandle&) = delete;
  ~DeviceHandle() { slpb_destroy(s); }
};

template <typename Scalar>
class Problem {
 public:
  Problem() : m_pool{&detail::pool()} {}
  /// OCP derives from Problem and is handed around as a Problem.
  virtual ~Problem() = default;

  /// Creates a decision variable in the optimization problem.
  [[nodiscard]] Variable<Scalar> decision_variable() {
    m_decision_variables.emplace_back();
    return m_decision_variables.back();
  }
  /// Creates a matrix of decision variables (appended row by row).
  [[nodiscard]] VariableMatrix<Scalar> decision_variable(int rows,
                                                         int cols = 1) {
    m_decision_variables.reserve(m_decision_variables.size() + rows * cols);
    VariableMatrix<Scalar> vars{detail::empty, rows, cols};
    for (int row = 0; row < rows; ++row) {
      for (int col = 0; col < cols; ++col) {
        m_decision_variables.emplace_back();
        vars(row, col) = m_decision_variables.back();
      }
    }
    return vars;
  }
  /// Creates a symmetric matrix of decision variables (lower triangle stored).
  [[nodiscard]] VariableMatrix<Scalar> symmetric_decision_variable(int rows) {
    VariableMatrix<Scalar> vars{detail::empty, rows, rows};
    for (int row = 0; row < rows; ++row) {
      for (int col = 0; col <= row; ++col) {
        m_decision_variables.emplace_back();
        vars(row, col) = m_decision_variables.back();
        vars(col, row) = m_decision_variables.back();
      }
    }
    return vars;
  }

  void minimize(const Variable<Scalar>& cost) { m_f = cost; }
  void maximize(const Variable<Scalar>& objective) { m_f = -objective; }

  void subject_to(const EqualityConstraints<Scalar>& constraint) {
    m_equality_constraints.insert(m_equality_constraints.end(),
                                  constraint.constraints.begin(),
                                  constraint.constraints.end());
  }
  void subject_to(const InequalityConstraints<Scalar>& constraint) {
    m_inequality_constraints.insert(m_inequality_constraints.end(),
                                    constraint.constraints.begin(),
                                    constraint.constraints.end());
  }

  ExpressionType cost_function_type() const {
    return m_f ? m_f->type() : ExpressionType::NONE;
  }
  ExpressionType equality_constraint_type() const {
    return max_type(m_equality_constraints);
  }
  ExpressionType inequality_constraint_type() const {
    return max_type(m_inequality_constraints);
  }

  /// Solves the optimization problem; the solution is stored in the original
  /// variables. With `spy` the sparsity patterns of H, A_e and A_i are
  /// written to H.spy / A_e.spy / A_i.spy at every iteration.
  ExitStatus solve(const Options& options = Options{}, bool spy = false) {
    DeviceOptions dev_options;
    dev_options.spy = spy;
    return solve(options, dev_options);
  }

  ExitStatus solve(const Options& options, const DeviceOptions& dev_options) {
    // Expression handles are ids into the CREATING thread's pool: a Problem
    // cannot be solved from another thread (the reference's pointers could).
    if (&detail::pool() != m_pool) {
      throw std::logic_error(
          "slp::Problem::solve called on another thread than the one that "
          "built the problem: expression handles are thread-local");
    }
    const auto t0 = std::chrono::steady_clock::now();
    ExitStatus status;
    {
      // the gradient trees, the Lagrangian and (if entered) the feasibility
      // restoration problem built below are dropped again before this returns:
      // give their nodes back to the pool, so that re-solving a long-lived
      // Problem (MPC, warm starts) does not grow it
      detail::PoolScope reclaim{detail::pool()};
      status = solve_impl(options, dev_options);
    }
    // everything solve_impl owned (graphs, flattened tape, device handle) has
    // been released by now: the remainder is teardown
    double accounted = 0.0;
    for (int i = 0; i < 8; ++i) accounted += m_phase[i];
    m_phase[8] = std::chrono::duration<double>(
                     std::chrono::steady_clock::now() - t0)
                     .count() -
                 accounted;
    return status;
  }

 private:
  ExitStatus solve_impl(const Options& options,
                        const DeviceOptions& dev_options) {
    m_trace = SolveTrace{};
    m_trace.keep_iterates = dev_options.keep_iterates;
    m_trace.flush_l2 = dev_options.flush_l2;

    const int n = static_cast<int>(m_decision_variables.size());
    const int me = static_cast<int>(m_equality_constraints.size());
    const int mi = static_cast<int>(m_inequality_constraints.size());

    // Initial value column vector (:288-291)
    std::vector<Scalar> x(n);
    for (int i = 0; i < n; ++i) x[i] = m_decision_variables[i].value();

    // Nothing to do for an empty or constant problem (:299-313)
    if (cost_function_type() <= ExpressionType::CONSTANT &&
        equality_constraint_type() <= ExpressionType::CONSTANT &&
        inequality_constraint_type() <= ExpressionType::CONSTANT) {
      return ExitStatus::SUCCESS;
    }

    std::vector<std::function<bool(const IterationInfo<Scalar>&)>> callbacks;
    for (const auto& cb : m_iteration_callbacks) callbacks.push_back(cb);
    for (const auto& cb : m_persistent_iteration_callbacks) {
      callbacks.push_back(cb);
    }

    // Sparsity pattern files (the reference registers the same callback)
    std::unique_ptr<Spy<Scalar>> H_spy, A_e_spy, A_i_spy;
    if (dev_options.spy) {
      H_spy = std::make_unique<Spy<Scalar>>("H.spy", "Hessian",
                                            "Decision variables",
                                            "Decision variables", n, n);
      if (me > 0) {
        A_e_spy = std::make_unique<Spy<Scalar>>(
            "A_e.spy", "Equality constraint Jacobian", "Constraints",
            "Decision variables", me, n);
      }
      if (mi > 0) {
        A_i_spy = std::make_unique<Spy<Scalar>>(
            "A_i.spy", "Inequality constraint Jacobian", "Constraints",
            "Decision variables", mi, n);
      }
      callbacks.push_back([&](const IterationInfo<Scalar>& info) -> bool {
        // restoration iterations hand in the enlarged problem's matrices
        if (info.H.rows() != n) return false;
        H_spy->add(info.H);
        if (A_e_spy) A_e_spy->add(info.A_e);
        if (A_i_spy) A_i_spy->add(info.A_i);
        return false;
      });
    }

    if (options.diagnostics) {
      const char* solver_name =
          me == 0 && mi == 0 ? "Newton" : (mi == 0 ? "SQP" : "IPM");
      std::printf("\nInvoking %s solver (B200 device path)\n\n", solver_name);
      std::printf("Number of decision variables: %d\n", n);
      std::printf("Number of equality constraints: %d\n", me);
      std::printf("Number of inequality constraints: %d\n\n", mi);
    }

    m_phase.fill(0.0);
    auto tick = std::chrono::steady_clock::now();
    auto lap = [&](int which) {
      const auto now = std::chrono::steady_clock::now();
      m_phase[which] += std::chrono::duration<double>(now - tick).count();
      tick = now;
    };

    // Autodiff setup (:517-560)
    auto graphs = build_graphs();
    lap(0);

    // Conflicting bounds (:597-606)
    if (has_conflicting_bounds(*graphs->A_i)) {
      return ExitStatus::GLOBALLY_INFEASIBLE;
    }

    // Hand the graphs to the device
    detail::FlatProblem fp = graphs->flatten();
    lap(1);

    DeviceHandle handle{dev_options.device};
    slpb_solver* dev = handle.s;
    if (dev_options.factor_arithmetic >= 0) {
      SLP_DEVICE_CALL(dev, slpb_set_factor_arithmetic(
                               dev, dev_options.factor_arithmetic));
    }
    if (dev_options.world > 1) {
      SLP_DEVICE_CALL(dev, slpb_comm_init(dev, dev_options.rank,
                                          dev_options.world,
                                          dev_options.nccl_unique_id.data()));
    }
    lap(2);
    upload(dev, fp);
    lap(3);

    // Initial iterate: s = 1, y = 0, z = 1 (interior_point.hpp:74-80)
    std::vector<Scalar> s(mi, Scalar(1)), y(me, Scalar(0)), z(mi, Scalar(1));
    SLP_DEVICE_CALL(dev,
                    slpb_set_iterate(dev, x.data(), s.data(), y.data(), z.data()));

    // Problem scaling from g(x₀), A_e(x₀), A_i(x₀) (:615-616;
    // problem_scaling.hpp:100-107)
    DeviceProblemInfo info{n, me, mi, 1.0, dev_options.world};
    std::vector<Scalar> d_ce, d_ci;
    {
      slpb_point_info pi{};
      SLP_DEVICE_CALL(dev, slpb_eval_current(dev, 1, &pi));
      std::vector<Scalar> gv(n);
      SLP_DEVICE_CALL(dev, slpb_download(dev, SLPB_ARR_G, gv.data()));
      Scalar g_inf(0);
      for (Scalar v : gv) g_inf = std::max(g_inf, std::abs(v));
      constexpr Scalar g_max(100);
      info.scaling_f = std::min(Scalar(1), g_max / g_inf);
      auto row_scales = [&](int which_pattern, int which_values, int rows) {
        std::vector<Scalar> norms(rows, Scalar(0));
        int32_t r = 0, c = 0;
        int64_t nnz = 0;
        SLP_DEVICE_CALL(dev, slpb_pattern(dev, which_pattern, &r, &c, &nnz,
                                          nullptr, nullptr));
        std::vector<int32_t> outer(c + 1), inner(nnz);
        std::vector<Scalar> values(nnz);
        SLP_DEVICE_CALL(dev, slpb_pattern(dev, which_pattern, &r, &c, &nnz,
                                          outer.data(), inner.data()));
        if (nnz > 0) {
          SLP_DEVICE_CALL(dev, slpb_download(dev, which_values, values.data()));
        }
        for (int64_t k = 0; k < nnz; ++k) {
          norms[inner[k]] = std::max(norms[inner[k]], std::abs(values[k]));
        }
        for (auto& v : norms) v = std::min(g_max / v, Scalar(1));
        return norms;
      };
      d_ce = row_scales(SLPB_OUT_A_E, SLPB_ARR_A_E_VAL, me);
      d_ci = row_scales(SLPB_OUT_A_I, SLPB_ARR_A_I_VAL, mi);
      SLP_DEVICE_CALL(dev, slpb_set_scaling(dev, info.scaling_f, d_ce.data(),
                                            d_ci.data()));
    }

    lap(4);
    slpb_symbolic_stats sym{};
    SLP_DEVICE_CALL(dev, slpb_analyze(dev, dev_options.ordering,
                                      dev_options.permutation.empty()
                                          ? nullptr
                                          : dev_options.permutation.data(),
                                      &sym));
    m_symbolic = sym;
    // one start of a slp::multistart: batch the linear algebra with the other
    // starts (refused — and then solved alone — if the pattern differs)
    if (detail::tls_multistart != nullptr && detail::tls_multistart->group &&
        !detail::tls_multistart_answered && dev_options.world == 1) {
      detail::tls_multistart_answered = true;
      (void)slpb_group_join(detail::tls_multistart->group, dev);
    }
    lap(5);

    // Interior-point method (:663-668; overload 1, interior_point.hpp:74-86)
    Scalar mu = Scalar(0.1) * info.scaling_f;
    int iterations = 0;
    const RestorationHook restoration =
        [&](double mu_outer, int& iters,
            const RestorationAcceptTest& accept) -> ExitStatus {
      return feasibility_restoration(dev, info, d_ce, d_ci, options, callbacks,
                                     mu_outer, iters, accept, dev_options);
    };
    // Newton without constraints (:335), SQP with equality constraints only
    // (:403), else the interior-point method (:512): one loop, three modes.
    SolverKind kind = SolverKind::IPM;
    if (me == 0 && mi == 0) {
      kind = SolverKind::NEWTON;
    } else if (mi == 0) {
      kind = SolverKind::SQP;
    }
    m_solver_kind = kind;
    ExitStatus status = interior_point<Scalar>(
        dev, info, std::span{callbacks}, options, false, mu, iterations,
        &m_trace, &restoration, 0.0, kind);

    lap(6);
    // Write the solution back into the Variables (:676)
    SLP_DEVICE_CALL(dev, slpb_get_iterate(dev, x.data(), s.data(), y.data(),
                                          z.data()));
    for (int i = 0; i < n; ++i) m_decision_variables[i].set_value(x[i]);
    m_last_s = std::move(s);
    m_last_y = std::move(y);
    m_last_z = std::move(z);
    slpb_get_counters(dev, &m_counters);
    slpb_get_timers(dev, &m_timers);
    slpb_get_comm_stats(dev, &m_comm_stats);
    lap(7);
    if (options.diagnostics) {
      if (iterations > 0) print_bottom_iteration_diagnostics();
      const std::string_view exit_name = to_string(status);
      std::printf("\nExit: %.*s after %d iterations\n\n",
                  static_cast<int>(exit_name.size()), exit_name.data(),
                  iterations);
      static constexpr const char* kHost[8] = {
          "autodiff setup", "flatten graphs", "device handle",
          "upload + compile", "problem scaling", "symbolic analysis",
          "Newton loop", "write-back"};
      std::vector<TimingRow> host_rows, device_rows;
      for (int i = 0; i < 8; ++i) {
        host_rows.push_back({kHost[i], 1e3 * m_phase[i], 1});
      }
      print_timing_table("host phases of solve()", host_rows);
      static constexpr const char* kDevice[5] = {
          "re-linearisation sweep", "value sweep (trial points)",
          "KKT assembly", "LDLT factorisation", "triangular solves"};
      for (int i = 0; i < 5; ++i) {
        // the device timers sample one run in eight: scale to all runs
        const double mean = m_timers.count[i] > 0
                                ? m_timers.total_ms[i] / double(m_timers.count[i])
                                : 0.0;
        device_rows.push_back({kDevice[i], mean * double(m_timers.launches[i]),
                               m_timers.launches[i]});
      }
      print_timing_table("device time per kernel group (sampled)", device_rows);
    }
    if (std::getenv("SLPB_DESTROY_TIMING")) {
      // where the teardown goes (development aid): release the big owners one
      // by one instead of at scope exit
      auto t = std::chrono::steady_clock::now();
      auto mark = [&](const char* what) {
        const auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[slp teardown] %-18s %8.2f ms\n", what,
                     std::chrono::duration<double, std::milli>(now - t).count());
        t = now;
      };
      fp = detail::FlatProblem{};
      mark("flat problem");
      graphs.reset();
      mark("graphs");
      slpb_destroy(handle.s);
      handle.s = nullptr;
      mark("device handle");
    }
    return status;
  }

 public:

  /// Adds a callback called at the beginning of each solver iteration; a
  /// void-returning callback never stops the solve.
  template <typename F>
    requires requires(F callback, const IterationInfo<Scalar>& info) {
      { callback(info) } -> std::same_as<void>;
    }
  void add_callback(F&& callback) {
    m_iteration_callbacks.emplace_back(
        [callback = std::forward<F>(callback)](
            const IterationInfo<Scalar>& info) {
          callback(info);
          return false;
        });
  }
  /// A bool-returning callback stops the solve by returning true.
  template <typename F>
    requires requires(F callback, const IterationInfo<Scalar>& info) {
      { callback(info) } -> std::same_as<bool>;
    }
  void add_callback(F&& callback) {
    m_iteration_callbacks.emplace_back(std::forward<F>(callback));
  }
  void clear_callbacks() { m_iteration_callbacks.clear(); }
  /// Persistent callbacks survive clear_callbacks() and run last.
  template <typename F>
    requires requires(F callback, const IterationInfo<Scalar>& info) {
      { callback(info) } -> std::same_as<bool>;
    }
  void add_persistent_callback(F&& callback) {
    m_persistent_iteration_callbacks.emplace_back(std::forward<F>(callback));
  }

  // ---- introspection of the last solve (not in the reference) -------------
  const SolveTrace& last_trace() const { return m_trace; }
  const slpb_symbolic_stats& last_symbolic_stats() const { return m_symbolic; }
  /// Which driver the last solve() took (Newton / SQP / IPM).
  SolverKind last_solver_kind() const { return m_solver_kind; }
  const slpb_counters& last_counters() const { return m_counters; }
  const slpb_timers& last_timers() const { return m_timers; }
  const slpb_comm_stats& last_comm_stats() const { return m_comm_stats; }
  /// Host seconds of the phases of the last solve(): build_graphs, flatten,
  /// device_create, upload+compile, scaling, analyze, newton loop, write-back,
  /// teardown (release of graphs, tape and device handle).
  const std::array<double, 9>& last_phase_seconds() const { return m_phase; }
  const std::vector<Scalar>& last_s() const { return m_last_s; }
  const std::vector<Scalar>& last_y() const { return m_last_y; }
  const std::vector<Scalar>& last_z() const { return m_last_z; }
  std::vector<Variable<Scalar>>& decision_variables() {
    return m_decision_variables;
  }
  const std::vector<Variable<Scalar>>& equality_constraints() const {
    return m_equality_constraints;
  }
  const std::vector<Variable<Scalar>>& inequality_constraints() const {
    return m_inequality_constraints;
  }
  const std::optional<Variable<Scalar>>& cost() const { return m_f; }

  /// The autodiff objects Problem::solve sets up for the interior-point
  /// method, in the reference's construction order (problem.hpp:517-560).
  struct Graphs {
    VariableMatrix<Scalar> x_ad, c_e_ad, c_i_ad, y_ad, z_ad;
    Variable<Scalar> f{nullptr};
    std::unique_ptr<Gradient<Scalar>> g;
    std::unique_ptr<Hessian<Scalar, Lower>> H_f, H_c;
    std::unique_ptr<Jacobian<Scalar>> A_e, A_i;

    detail::FlatProblem flatten() const {
      detail::Flattener flat;
      flat.set_leaves(x_ad, y_ad, z_ad);
      flat.add_values(SLPB_OUT_F, VariableMatrix<Scalar>{f});
      flat.add_jacobian(SLPB_OUT_G, g->jacobian());
      flat.add_jacobian(SLPB_OUT_H_F, *H_f);
      flat.add_jacobian(SLPB_OUT_H_C, *H_c);
      flat.add_values(SLPB_OUT_C_E, c_e_ad);
      flat.add_jacobian(SLPB_OUT_A_E, *A_e);
      flat.add_values(SLPB_OUT_C_I, c_i_ad);
      flat.add_jacobian(SLPB_OUT_A_I, *A_i);
      return flat.finish();
    }
  };

  std::unique_ptr<Graphs> build_graphs() {
    return build_graphs_from(m_decision_variables,
                             m_f.value_or(Variable<Scalar>{Scalar(0)}),
                             m_equality_constraints, m_inequality_constraints);
  }

  /// The same construction over explicit lists (the feasibility-restoration
  /// problem reuses the original constraint expressions over an enlarged
  /// variable vector).
  static std::unique_ptr<Graphs> build_graphs_from(
      const std::vector<Variable<Scalar>>& xs, const Variable<Scalar>& f,
      const std::vector<Variable<Scalar>>& c_e,
      const std::vector<Variable<Scalar>>& c_i) {
    const int me = static_cast<int>(c_e.size());
    const int mi = static_cast<int>(c_i.size());
    auto G = std::make_unique<Graphs>();
    G->x_ad = VariableMatrix<Scalar>{
        std::span<const Variable<Scalar>>{xs.data(), xs.size()}};
    G->f = f;
    G->c_e_ad = VariableMatrix<Scalar>{
        std::span<const Variable<Scalar>>{c_e.data(), c_e.size()}};
    G->c_i_ad = VariableMatrix<Scalar>{
        std::span<const Variable<Scalar>>{c_i.data(), c_i.size()}};
    G->y_ad = VariableMatrix<Scalar>(me);
    G->z_ad = VariableMatrix<Scalar>(mi);
    G->g = std::make_unique<Gradient<Scalar>>(G->f, G->x_ad);
    G->H_f = std::make_unique<Hessian<Scalar, Lower>>(G->f, G->x_ad);
    // −yᵀc_e − zᵀc_i as 1x1 matrix products: left-deep chains (:547-548)
    Variable<Scalar> lagrangian_c{VariableMatrix<Scalar>{
        -G->y_ad.T() * G->c_e_ad - G->z_ad.T() * G->c_i_ad}};
    G->H_c = std::make_unique<Hessian<Scalar, Lower>>(lagrangian_c, G->x_ad);
    G->A_e = std::make_unique<Jacobian<Scalar>>(G->c_e_ad, G->x_ad);
    G->A_i = std::make_unique<Jacobian<Scalar>>(G->c_i_ad, G->x_ad);
    return G;
  }

  /// Uploads a flattened problem and compiles it on the device.
  static void upload(slpb_solver* dev, const detail::FlatProblem& fp) {
    SLP_DEVICE_CALL(
        dev, slpb_upload_tape(dev, fp.n_nodes(), fp.op.data(), fp.lhs.data(),
                              fp.rhs.data(), fp.val.data(),
                              static_cast<int32_t>(fp.leaf_x.size()),
                              fp.leaf_x.data(),
                              static_cast<int32_t>(fp.leaf_y.size()),
                              fp.leaf_y.data(),
                              static_cast<int32_t>(fp.leaf_z.size()),
                              fp.leaf_z.data()));
    for (int which = 0; which < SLPB_OUT_COUNT; ++which) {
      const slpb_rowset view = fp.rows[which].view();
      SLP_DEVICE_CALL(dev, slpb_upload_rows(dev, which, &view,
                                            fp.rows[which].const_val.data()));
    }
    SLP_DEVICE_CALL(dev, slpb_finalize(dev));
  }

 private:
  /// Feasibility restoration for the interior-point method
  /// (solver/util/feasibility_restoration.hpp:346-628). The reference wraps the
  /// matrix callbacks into an enlarged problem over x̃ = [x, p_e, n_e, p_i, n_i]:
  ///
  ///   min ρΣ(p + n) + ζ/2 (x − x_r)ᵀD_r(x − x_r)
  ///   s.t. c_e(x) − p_e + n_e = 0,  c_i(x) − p_i + n_i ≥ 0,  p, n ≥ 0
  ///
  /// and re-enters interior_point on it. Here the enlarged problem is written
  /// with the DSL over the ORIGINAL constraint expressions, compiled for a
  /// second device handle (Hessian of the constraints ignored, as the
  /// reference's H callback effectively does, :485-495) and solved by the same
  /// device loop with in_feasibility_restoration = true.
  ExitStatus feasibility_restoration(
      slpb_solver* dev, const DeviceProblemInfo& info,
      const std::vector<Scalar>& d_ce, const std::vector<Scalar>& d_ci,
      const Options& options,
      std::vector<std::function<bool(const IterationInfo<Scalar>&)>>& callbacks,
      Scalar mu, int& iterations, const RestorationAcceptTest& accept,
      const DeviceOptions& dev_options) {
    const int n = info.num_decision_variables;
    const int me = info.num_equality_constraints;
    const int mi = info.num_inequality_constraints;
    constexpr Scalar rho(1e3);

    std::vector<Scalar> x(n), s(mi), y(me), z(mi), c_e(me), c_i(mi);
    SLP_DEVICE_CALL(dev, slpb_get_iterate(dev, x.data(), s.data(), y.data(),
                                          z.data()));
    if (me > 0) SLP_DEVICE_CALL(dev, slpb_download(dev, SLPB_ARR_C_E, c_e.data()));
    if (mi > 0) SLP_DEVICE_CALL(dev, slpb_download(dev, SLPB_ARR_C_I, c_i.data()));

    Scalar fr_mu = mu;
    for (Scalar v : c_e) fr_mu = std::max(fr_mu, std::abs(v));
    std::vector<Scalar> cis(mi);
    for (int i = 0; i < mi; ++i) {
      cis[i] = c_i[i] - s[i];
      fr_mu = std::max(fr_mu, std::abs(cis[i]));
    }
    const Scalar zeta = std::sqrt(fr_mu);

    // closed-form initial p, n (:49-100)
    auto compute_p_n = [&](const std::vector<Scalar>& c, std::vector<Scalar>& p,
                           std::vector<Scalar>& nn) {
      p.resize(c.size());
      nn.resize(c.size());
      for (size_t r = 0; r < c.size(); ++r) {
        const Scalar a = rho;
        const Scalar b = rho * c[r] - fr_mu;
        const Scalar cc = -fr_mu * c[r] / Scalar(2);
        nn[r] = (-b + std::sqrt(b * b - Scalar(4) * a * cc)) / (Scalar(2) * a);
        p[r] = c[r] + nn[r];
      }
    };
    std::vector<Scalar> p_e_0, n_e_0, p_i_0, n_i_0;
    compute_p_n(c_e, p_e_0, n_e_0);
    compute_p_n(cis, p_i_0, n_i_0);

    // ---- the enlarged problem, over the original constraint expressions ----
    const int extra = 2 * me + 2 * mi;
    std::vector<Variable<Scalar>> xs = m_decision_variables;
    std::vector<Variable<Scalar>> p_e(me), n_e(me), p_i(mi), n_i(mi);
    for (auto* blk : {&p_e, &n_e, &p_i, &n_i}) {
      xs.insert(xs.end(), blk->begin(), blk->end());
    }
    // cost = ρΣ(p + n) + ζ/2 (x − x_r)ᵀD_r(x − x_r), written as ONE flat sum of
    // small terms (the factors ρ and ζ/2 inside each term) so that the device
    // compiler can split the row into independent clusters instead of one
    // 10⁵-node chain.
    Variable<Scalar> cost{Scalar(0)};
    for (int i = n; i < n + extra; ++i) cost = cost + rho * xs[i];
    for (int i = 0; i < n; ++i) {
      const Scalar D_r = std::min(Scalar(1) / (x[i] * x[i]), Scalar(1));
      Variable<Scalar> d = xs[i] - x[i];
      cost = cost + (zeta / Scalar(2)) * (d * (D_r * d));
    }

    // rows are scaled as a whole on the device (d_c ⊙ row); the reference adds
    // −p + n to the already scaled d_c ⊙ c(x), so p and n enter divided by d_c
    auto relax = [](const Variable<Scalar>& c, const Variable<Scalar>& p,
                    const Variable<Scalar>& nn, Scalar d) {
      if (d == Scalar(1)) return c - p + nn;
      return c - p / d + nn / d;
    };
    std::vector<Variable<Scalar>> ce_fr, ci_fr;
    ce_fr.reserve(me);
    ci_fr.reserve(mi + extra);
    for (int j = 0; j < me; ++j) {
      ce_fr.push_back(relax(m_equality_constraints[j], p_e[j], n_e[j], d_ce[j]));
    }
    for (int i = 0; i < mi; ++i) {
      ci_fr.push_back(
          relax(m_inequality_constraints[i], p_i[i], n_i[i], d_ci[i]));
    }
    for (int i = n; i < n + extra; ++i) ci_fr.push_back(xs[i]);

    auto graphs = build_graphs_from(xs, cost, ce_fr, ci_fr);
    detail::FlatProblem fp = graphs->flatten();

    DeviceHandle inner{dev_options.device};
    SLP_DEVICE_CALL(inner.s, slpb_set_ignore_constraint_hessian(inner.s, 1));
    if (dev_options.factor_arithmetic >= 0) {
      SLP_DEVICE_CALL(inner.s, slpb_set_factor_arithmetic(
                                   inner.s, dev_options.factor_arithmetic));
    }
    upload(inner.s, fp);

    // initial iterate (:408-424): perfect complementarity with the slacks
    std::vector<Scalar> fr_x(x), fr_s(s), fr_y(me, Scalar(0)), fr_z;
    fr_s.resize(mi + extra, Scalar(1));
    fr_z.reserve(mi + extra);
    for (const auto* v : {&p_e_0, &n_e_0, &p_i_0, &n_i_0}) {
      fr_x.insert(fr_x.end(), v->begin(), v->end());
    }
    for (const auto* v : {&s, &p_e_0, &n_e_0, &p_i_0, &n_i_0}) {
      for (Scalar e : *v) fr_z.push_back(fr_mu * (Scalar(1) / e));
    }
    SLP_DEVICE_CALL(inner.s, slpb_set_iterate(inner.s, fr_x.data(), fr_s.data(),
                                              fr_y.data(), fr_z.data()));
    std::vector<Scalar> fr_d_ci(d_ci);
    fr_d_ci.resize(mi + extra, Scalar(1));
    SLP_DEVICE_CALL(inner.s, slpb_set_scaling(inner.s, 1.0, d_ce.data(),
                                              fr_d_ci.data()));
    slpb_symbolic_stats sym{};
    SLP_DEVICE_CALL(inner.s, slpb_analyze(inner.s, SLPB_ORDER_NESTED_DISSECTION,
                                          nullptr, &sym));

    // user callbacks + the acceptance test (interior_point.hpp:733-756)
    std::vector<std::function<bool(const IterationInfo<Scalar>&)>> fr_callbacks{
        callbacks.begin(), callbacks.end()};
    fr_callbacks.emplace_back([&](const IterationInfo<Scalar>& it) {
      std::vector<double> trial_x(it.x.data(), it.x.data() + n);
      std::vector<double> trial_s(it.s.data(), it.s.data() + mi);
      return accept(trial_x, trial_s);
    });

    const DeviceProblemInfo fr_info{n + extra, me, mi + extra, 1.0};
    Scalar fr_mu_io = fr_mu;
    const ExitStatus status = interior_point<Scalar>(
        inner.s, fr_info, std::span{fr_callbacks}, options, true, fr_mu_io,
        iterations, &m_trace);

    SLP_DEVICE_CALL(inner.s, slpb_get_iterate(inner.s, fr_x.data(), fr_s.data(),
                                              nullptr, nullptr));
    std::copy(fr_x.begin(), fr_x.begin() + n, x.begin());
    std::copy(fr_s.begin(), fr_s.begin() + mi, s.begin());
    SLP_DEVICE_CALL(dev,
                    slpb_set_iterate(dev, x.data(), s.data(), y.data(), z.data()));
    {
      slpb_counters c{};
      slpb_get_counters(inner.s, &c);
      m_restoration_launches += c.kernel_launches;
    }

    if (status == ExitStatus::CALLBACK_REQUESTED_STOP) {
      // y, z by least squares at the restored point (:612-622)
      slpb_point_info pi{};
      SLP_DEVICE_CALL(dev, slpb_eval_current(dev, 1, &pi));
      slpb_factor_info fi{};
      SLP_DEVICE_CALL(dev, slpb_multiplier_estimate(dev, mu, &fi));
      if (fi.zero_pivot) return ExitStatus::FEASIBILITY_RESTORATION_FAILED;
      return ExitStatus::SUCCESS;
    } else if (status == ExitStatus::SUCCESS) {
      return ExitStatus::LOCALLY_INFEASIBLE;  // :623-624
    }
    return ExitStatus::FEASIBILITY_RESTORATION_FAILED;
  }

  static ExpressionType max_type(const std::vector<Variable<Scalar>>& v) {
    ExpressionType t = ExpressionType::NONE;
    for (const auto& e : v) t = std::max(t, e.type());
    return t;
  }

  /// solver/util/bounds.hpp:54-179, reduced to what solve() consumes: whether
  /// two bound constraints on one variable contradict each other. A bound is a
  /// LINEAR inequality whose gradient has a single structural entry.
  bool has_conflicting_bounds(const Jacobian<Scalar>& A_i) {
    const int mi = static_cast<int>(m_inequality_constraints.size());
    std::vector<int> count(mi, 0), col(mi, -1);
    std::vector<Scalar> coeff(mi, Scalar(0));
    for (const auto& t : A_i.cached_triplets()) {
      ++count[t.row];
      col[t.row] = t.col;
      coeff[t.row] = t.value;
    }
    const size_t n = m_decision_variables.size();
    std::vector<std::pair<Scalar, Scalar>> bnd(
        n, {-std::numeric_limits<Scalar>::infinity(),
            std::numeric_limits<Scalar>::infinity()});
    bool conflict = false;
    for (int ci = 0; ci < mi; ++ci) {
      if (m_inequality_constraints[ci].type() != ExpressionType::LINEAR) {
        continue;
      }
      if (count[ci] != 1) continue;
      auto& var = m_decision_variables[col[ci]];
      const Scalar var_value = var.value();
      Scalar constant;
      if (var_value != Scalar(0)) {
        var.set_value(Scalar(0));
        constant = m_inequality_constraints[ci].value();
        var.set_value(var_value);
      } else {
        constant = m_inequality_constraints[ci].value();
      }
      auto& [lower, upper] = bnd[col[ci]];
      const Scalar detected = -constant / coeff[ci];
      if (coeff[ci] < Scalar(0) && detected < upper) {
        upper = detected;
      } else if (coeff[ci] > Scalar(0) && detected > lower) {
        lower = detected;
      }
      if (lower > upper) conflict = true;
    }
    return conflict;
  }

  std::vector<Variable<Scalar>> m_decision_variables;
  std::optional<Variable<Scalar>> m_f;
  std::vector<Variable<Scalar>> m_equality_constraints;
  std::vector<Variable<Scalar>> m_inequality_constraints;
  std::vector<std::function<bool(const IterationInfo<Scalar>&)>>
      m_iteration_callbacks;
  std::vector<std::function<bool(const IterationInfo<Scalar>&)>>
      m_persistent_iteration_callbacks;

  SolveTrace m_trace;
  slpb_symbolic_stats m_symbolic{};
  SolverKind m_solver_kind = SolverKind::IPM;
  detail::ExpressionPool* m_pool;  ///< pool of the thread that built the problem
  slpb_counters m_counters{};
  slpb_timers m_timers{};
  slpb_comm_stats m_comm_stats{};
  std::array<double, 9> m_phase{};
  int64_t m_restoration_launches = 0;
  std::vector<Scalar> m_last_s, m_last_y, m_last_z;
};

}  // namespace slp
