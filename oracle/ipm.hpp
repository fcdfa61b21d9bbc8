// ORACLE — TEST INFRASTRUCTURE ONLY (see expr.hpp header).
//
// CPU restatement of the reference's interior-point solver and helpers:
//   optimization/solver/interior_point.hpp:63-866 (both overloads, the loop)
//   solver/util/filter.hpp:53-57,87-172      solver/util/kkt_error.hpp:92-251
//   solver/util/fraction_to_the_boundary_rule.hpp:19-43
//   solver/util/is_locally_infeasible.hpp:17-60
//   solver/util/problem_scaling.hpp:21-115
//   solver/util/regularized_ldlt.hpp, sparse_regularized_ldlt.hpp:64-152,
//   dense_regularized_ldlt.hpp:59-136, inertia.hpp
//   solver/util/feasibility_restoration.hpp:49-100,346-628
//   solver/util/lagrange_multiplier_estimate.hpp:55-131
//   solver/options.hpp, exit_status.hpp, iteration_info.hpp
// Iterate-level parity is UNPINNED in the reference (no test checks an
// intermediate iterate) — this file is validated through the
// reference's solution-level known-answer tests (tests/test_oracle_*.py).
#pragma once

#include <algorithm>
#include <chrono>
#include <cmath>
#include <functional>
#include <limits>
#include <tuple>
#include <vector>

#include "ldlt.hpp"
#include "sparse.hpp"

namespace orc {

enum class ExitStatus : int8_t {  // exit_status.hpp:13-44
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

struct Options {  // options.hpp:13-38
  double tolerance = 1e-8;
  int max_iterations = 5000;
  double timeout = std::numeric_limits<double>::infinity();  // seconds
  bool feasible_ipm = false;
  bool diagnostics = false;
};

struct ProblemScaling {  // problem_scaling.hpp
  double f = 1.0;
  Vec c_e, c_i;
  ProblemScaling() = default;
  ProblemScaling(double f_, Vec ce, Vec ci)
      : f{f_}, c_e{std::move(ce)}, c_i{std::move(ci)} {}
  /// :100-107 — d_f = min(1, 100/‖g‖∞), d_c[j] = min(1, 100/‖row j‖∞).
  ProblemScaling(const Vec& g, const Csc& A_e, const Csc& A_i) {
    constexpr double g_max = 100.0;
    f = std::min(1.0, g_max / norm_inf(g));
    auto rows = [&](const Csc& A) {
      Vec n = A.row_inf_norms();
      for (auto& v : n) v = std::min(g_max / v, 1.0);
      return n;
    };
    c_e = rows(A_e);
    c_i = rows(A_i);
  }
  bool is_identity() const {  // :112-114
    return f == 1.0 && c_e.empty() && c_i.empty();
  }
};

struct IterationInfo {  // iteration_info.hpp:13-41
  int iteration;
  const Vec& x;
  const Vec& s;
  const Vec& y;
  const Vec& z;
  const Vec& g;  // dense here; SparseVector in the reference
  const Csc& H;
  const Csc& A_e;
  const Csc& A_i;
};
using IterationCallback = std::function<bool(const IterationInfo&)>;

struct IpmCallbacks {  // interior_point_matrix_callbacks.hpp:18-250
  int num_decision_variables = 0;
  int num_equality_constraints = 0;
  int num_inequality_constraints = 0;
  std::function<double(const Vec&)> f;
  std::function<Vec(const Vec&)> g;
  std::function<Csc(const Vec&, const Vec&, const Vec&)> H;
  std::function<Csc(const Vec&, const Vec&, const Vec&)> H_c;
  std::function<Vec(const Vec&)> c_e;
  std::function<Csc(const Vec&)> A_e;
  std::function<Vec(const Vec&)> c_i;
  std::function<Csc(const Vec&)> A_i;
  ProblemScaling scaling;
};

/// One row per Newton iteration, recorded at the end of the loop body. Not in
/// the reference: this is the iterate dump the parity tests compare against.
struct TraceRow {
  int iteration;
  int type;  // 0 normal, 1 inside feasibility restoration
  double error, cost, infeasibility, complementarity, mu, delta, gamma;
  double alpha, alpha_max, alpha_z;
  int factorizations, solves, trials;
  double t_end = 0.0;  // seconds since interior_point() was entered
  Vec x, s, y, z;
};
struct Trace {
  bool keep_iterates = true;
  std::vector<TraceRow> rows;
  int total_factorizations = 0, total_solves = 0, total_trials = 0;
};

struct Inertia {  // inertia.hpp
  int positive = 0, negative = 0, zero = 0;
  Inertia() = default;
  Inertia(int p, int n, int z) : positive{p}, negative{n}, zero{z} {}
  explicit Inertia(const Vec& D) {
    constexpr double eps = std::numeric_limits<double>::epsilon();
    for (double e : D) {
      if (e > eps) {
        ++positive;
      } else if (e < -eps) {
        ++negative;
      } else {
        ++zero;
      }
    }
  }
  bool operator==(const Inertia&) const = default;
};

/// regularized_ldlt.hpp + sparse_/dense_regularized_ldlt.hpp.
class RegularizedLDLT {
 public:
  RegularizedLDLT(bool sparse, int n, int me, double gamma_min)
      : m_sparse{sparse}, m_n{n}, m_me{me}, m_gamma_min{gamma_min},
        m_ideal{n, me, 0} {}

  SimplicialLDLT& sparse_solver() { return m_sp; }
  int factorizations = 0;

  /// Returns true on success (info() == Success).
  bool compute(const Csc& lhs) {
    // lhs + reg(0,0): forces a full explicit diagonal so the pattern is stable
    // (sparse_regularized_ldlt.hpp:65-72).
    bool ok = factor(lhs, 0.0, 0.0, true);
    if (ok) {
      const Vec& D = m_D;
      bool far = std::all_of(D.begin(), D.end(),
                             [](double d) { return std::abs(d) >= 1e-4; });
      if (Inertia{D} == m_ideal && far) {
        m_prev_delta = 0.0;
        m_prev_gamma = 0.0;
        return m_ok = true;
      }
    }
    double delta =
        m_prev_delta == 0.0
            ? 1e-4
            : std::max(m_prev_delta / 2.0,
                       std::numeric_limits<double>::epsilon());
    double gamma = m_gamma_min;
    while (true) {
      ok = factor(lhs, delta, gamma, false);
      if (ok) {
        Inertia inertia{m_D};
        if (inertia == m_ideal) {
          m_prev_delta = delta;
          m_prev_gamma = gamma;
          return m_ok = true;
        } else if (inertia.zero > 0) {
          if (gamma == 0.0) {
            gamma = 1e-10;
          } else {
            delta *= 10.0;
            gamma *= 10.0;
          }
        } else if (inertia.negative > m_ideal.negative) {
          delta *= 10.0;
        } else if (inertia.positive > m_ideal.positive) {
          gamma = gamma == 0.0 ? 1e-10 : gamma * 10.0;
        }
      } else {
        delta *= 10.0;
        gamma = gamma == 0.0 ? 1e-10 : gamma * 10.0;
      }
      if (delta > 1e20 || gamma > 1e20) {
        m_prev_delta = delta;
        m_prev_gamma = gamma;
        return m_ok = false;
      }
    }
  }

  Vec solve(const Vec& rhs) const {
    return m_sparse ? m_sp.solve(rhs) : m_dn.solve(rhs);
  }
  bool ok() const { return m_ok; }
  double hessian_regularization() const { return m_prev_delta; }
  double constraint_jacobian_regularization() const { return m_prev_gamma; }

 private:
  bool factor(const Csc& lhs, double delta, double gamma, bool first) {
    ++factorizations;
    int dim = m_n + m_me;
    Vec reg(dim);
    for (int i = 0; i < m_n; ++i) reg[i] = delta;
    for (int i = 0; i < m_me; ++i) reg[m_n + i] = -gamma;
    Csc K = add(lhs, diag(reg));
    bool ok;
    if (m_sparse) {
      if (first && !m_sp.analyzed()) m_sp.analyze(K);
      ok = m_sp.factorize(K);
      if (ok) m_D = m_sp.vectorD();
    } else {
      std::vector<double> dense(size_t(dim) * dim, 0.0);
      for (int c = 0; c < dim; ++c) {
        for (int k = K.colptr[c]; k < K.colptr[c + 1]; ++k) {
          dense[size_t(K.rowidx[k]) * dim + c] = K.val[k];
        }
      }
      ok = m_dn.compute(std::move(dense), dim);
      if (ok) m_D = m_dn.vectorD();
    }
    return ok;
  }

  bool m_sparse;
  int m_n, m_me;
  double m_gamma_min;
  Inertia m_ideal;
  SimplicialLDLT m_sp;
  DenseLDLT m_dn;
  Vec m_D;
  bool m_ok = true;
  double m_prev_delta = 0.0, m_prev_gamma = 0.0;
};

struct FilterEntry {  // filter.hpp:20-63
  double cost = 0.0;
  double constraint_violation = 0.0;
  FilterEntry() = default;
  FilterEntry(double c, double v) : cost{c}, constraint_violation{v} {}
  FilterEntry(double f, const Vec& s, const Vec& c_e, const Vec& c_i,
              double mu) {
    double logsum = 0.0;
    for (double v : s) logsum += std::log(v);
    cost = f - mu * logsum;
    constraint_violation = norm1(c_e) + norm1(sub(c_i, s));
  }
  bool dominated_by(const FilterEntry& e) const {
    return e.cost <= cost && e.constraint_violation <= constraint_violation;
  }
};

class Filter {  // filter.hpp:70-212
 public:
  double min_constraint_violation;
  double max_constraint_violation;
  explicit Filter(double initial_violation = 0.0) {
    min_constraint_violation = 1e-4 * std::max(1.0, initial_violation);
    max_constraint_violation = 1e4 * std::max(1.0, initial_violation);
  }
  void reset() {
    m_filter.clear();
    m_last_rejection_due_to_filter = false;
  }
  bool try_add(const FilterEntry& cur, const FilterEntry& trial, double D_phi,
               double alpha) {
    if (!std::isfinite(trial.cost) ||
        trial.constraint_violation > max_constraint_violation) {
      return false;
    }
    constexpr double s_phi = 2.3, s_theta = 1.1;
    bool switching =
        D_phi < 0.0 && alpha * std::pow(-D_phi, s_phi) >
                           std::pow(cur.constraint_violation, s_theta);
    constexpr double eta_phi = 1e-8;
    bool armijo = trial.cost <= cur.cost + eta_phi * alpha * D_phi;
    double phi = std::pow(alpha, 1.5);
    bool sufficient =
        trial.cost <= cur.cost - phi * kGammaCost * cur.constraint_violation ||
        trial.constraint_violation <=
            (1.0 - phi * kGammaConstraint) * cur.constraint_violation;
    if (cur.constraint_violation <= min_constraint_violation && switching) {
      if (!armijo) {
        m_last_rejection_due_to_filter = false;
        return false;
      }
    } else if (!sufficient) {
      m_last_rejection_due_to_filter = false;
      return false;
    }
    if (in_filter(trial)) {
      m_last_rejection_due_to_filter = true;
      return false;
    }
    if (!switching || !armijo) {
      add(FilterEntry{
          cur.cost - phi * kGammaCost * cur.constraint_violation,
          (1.0 - phi * kGammaConstraint) * cur.constraint_violation});
    }
    return true;
  }
  bool last_rejection_due_to_filter() const {
    return m_last_rejection_due_to_filter;
  }

 private:
  static constexpr double kGammaCost = 1e-8;
  static constexpr double kGammaConstraint = 1e-5;
  std::vector<FilterEntry> m_filter;
  bool m_last_rejection_due_to_filter = false;
  void add(const FilterEntry& e) {
    std::erase_if(m_filter,
                  [&](const FilterEntry& x) { return x.dominated_by(e); });
    m_filter.push_back(e);
  }
  bool in_filter(const FilterEntry& e) const {
    return std::any_of(m_filter.begin(), m_filter.end(),
                       [&](const FilterEntry& x) { return e.dominated_by(x); });
  }
};

/// fraction_to_the_boundary_rule.hpp:19-43 (sequential scan, divide first).
inline double fraction_to_the_boundary_rule(const Vec& x, const Vec& p,
                                            double tau) {
  double alpha = 1.0;
  for (size_t i = 0; i < x.size(); ++i) {
    if (alpha * p[i] < -tau * x[i]) alpha = -tau / p[i] * x[i];
  }
  return alpha;
}

enum class KKTErrorType { INF_NORM_SCALED, ONE_NORM };

/// Which of the reference's three Newton-type drivers the loop below restates.
/// They share one structure (optimization/solver/{interior_point,sqp,newton}.hpp);
/// SQP is the loop without slacks, barrier and fraction-to-the-boundary rule
/// (sqp.hpp:91-596), Newton additionally drops the constraints, the
/// second-order correction and feasibility restoration (newton.hpp:50-290).
enum class SolverKind { IPM, SQP, NEWTON };

/// kkt_error.hpp:92-146.
inline double kkt_error(KKTErrorType T, const Vec& g, const Csc& A_e,
                        const Vec& c_e, const Csc& A_i, const Vec& c_i,
                        const Vec& s, const Vec& y, const Vec& z, double mu) {
  Vec r = g;
  Vec aty = A_e.mul_t(y);
  Vec atz = A_i.mul_t(z);
  for (size_t i = 0; i < r.size(); ++i) r[i] = r[i] - aty[i] - atz[i];
  Vec comp(s.size());
  for (size_t i = 0; i < s.size(); ++i) comp[i] = s[i] * z[i] - mu;
  Vec cis = sub(c_i, s);
  if (T == KKTErrorType::INF_NORM_SCALED) {
    constexpr double s_max = 100.0;
    double s_d = std::max(s_max, (norm1(y) + norm1(z)) /
                                     double(y.size() + z.size())) /
                 s_max;
    double s_c = std::max(s_max, norm1(z) / double(z.size())) / s_max;
    return std::max({norm_inf(r) / s_d, norm_inf(comp) / s_c, norm_inf(c_e),
                     norm_inf(cis)});
  }
  return norm1(r) + norm1(comp) + norm1(c_e) + norm1(cis);
}

/// kkt_error.hpp:216-251.
inline double unscaled_kkt_error(KKTErrorType T, const ProblemScaling& sc,
                                 const Vec& g, const Csc& A_e, const Vec& c_e,
                                 const Csc& A_i, const Vec& c_i, const Vec& s,
                                 const Vec& y, const Vec& z, double mu) {
  if (sc.is_identity()) {
    return kkt_error(T, g, A_e, c_e, A_i, c_i, s, y, z, mu);
  }
  const double inv_d_f = 1.0 / sc.f;
  Vec inv_ce(sc.c_e.size()), inv_ci(sc.c_i.size());
  for (size_t i = 0; i < inv_ce.size(); ++i) inv_ce[i] = 1.0 / sc.c_e[i];
  for (size_t i = 0; i < inv_ci.size(); ++i) inv_ci[i] = 1.0 / sc.c_i[i];
  Vec g_u(g.size());
  for (size_t i = 0; i < g.size(); ++i) g_u[i] = inv_d_f * g[i];
  Csc A_e_u = A_e.scale_rows(inv_ce);
  Csc A_i_u = A_i.scale_rows(inv_ci);
  Vec c_e_u(c_e.size()), y_u(y.size());
  for (size_t i = 0; i < c_e.size(); ++i) {
    c_e_u[i] = inv_ce[i] * c_e[i];
    y_u[i] = sc.c_e[i] * y[i] * inv_d_f;
  }
  Vec c_i_u(c_i.size()), s_u(s.size()), z_u(z.size());
  for (size_t i = 0; i < c_i.size(); ++i) {
    c_i_u[i] = inv_ci[i] * c_i[i];
    s_u[i] = inv_ci[i] * s[i];
    z_u[i] = sc.c_i[i] * z[i] * inv_d_f;
  }
  return kkt_error(T, g_u, A_e_u, c_e_u, A_i_u, c_i_u, s_u, y_u, z_u,
                   inv_d_f * mu);
}

/// is_locally_infeasible.hpp:17-60.
inline bool is_equality_locally_infeasible(const Csc& A_e, const Vec& c_e) {
  return A_e.rows > 0 && norm2(A_e.mul_t(c_e)) < 1e-6 && norm2(c_e) > 1e-2;
}
inline bool is_inequality_locally_infeasible(const Csc& A_i, const Vec& c_i) {
  if (A_i.rows > 0) {
    Vec plus(c_i.size());
    for (size_t i = 0; i < c_i.size(); ++i) plus[i] = std::min(c_i[i], 0.0);
    if (norm2(A_i.mul_t(plus)) < 1e-6 && norm2(plus) > 1e-6) return true;
  }
  return false;
}

/// How the oracle orders the sparse KKT factor. AMD restates Eigen's default;
/// CUSTOM lets a parity test hand in the permutation the CUDA path used.
struct LinearSolverConfig {
  SimplicialLDLT::Ordering ordering = SimplicialLDLT::Ordering::AMD;
  std::vector<int> permutation;
  int force_sparse = -1;  // -1: reference rule (:340-348), 0 dense, 1 sparse
};

inline ExitStatus interior_point(const IpmCallbacks& matrices,
                                 std::vector<IterationCallback>& callbacks,
                                 const Options& options,
                                 bool in_feasibility_restoration, Vec& x,
                                 Vec& s, Vec& y, Vec& z, double& mu,
                                 int& iterations, Trace* trace,
                                 const LinearSolverConfig& lin,
                                 SolverKind kind = SolverKind::IPM);

/// lagrange_multiplier_estimate.hpp:55-131.
inline std::pair<Vec, Vec> lagrange_multiplier_estimate(const Vec& g,
                                                        const Csc& A_e,
                                                        const Csc& A_i,
                                                        const Vec& s,
                                                        double mu) {
  // Â = [A_e 0; A_i −S]
  int n = A_e.cols, me = A_e.rows, mi = A_i.rows;
  Vec neg_s(s.size());
  for (size_t i = 0; i < s.size(); ++i) neg_s[i] = -s[i];
  Csc left = vstack(A_e, A_i);
  Csc A_hat{me + mi, n + mi};
  A_hat.colptr = left.colptr;
  A_hat.rowidx = left.rowidx;
  A_hat.val = left.val;
  A_hat.colptr.resize(n + mi + 1);
  for (int j = 0; j < mi; ++j) {
    A_hat.rowidx.push_back(me + j);
    A_hat.val.push_back(neg_s[j]);
    A_hat.colptr[n + j + 1] = A_hat.nnz();
  }
  Csc lhs = matmul(A_hat, A_hat.transpose());
  Vec rhs_temp(n + mi);
  for (int i = 0; i < n; ++i) rhs_temp[i] = g[i];
  for (int i = 0; i < mi; ++i) rhs_temp[n + i] = -mu;
  Vec rhs = A_hat.mul(rhs_temp);
  SimplicialLDLT est;
  Csc lhs_lower = lhs.lower();
  est.analyze(lhs_lower);
  est.factorize(lhs_lower);
  Vec sol = est.solve(rhs);
  Vec yv(sol.begin(), sol.begin() + me);
  Vec zv(sol.begin() + me, sol.end());
  for (int r = 0; r < mi; ++r) {
    constexpr double kappa = 1e10;
    zv[r] = std::clamp(zv[r], 1.0 / kappa * mu / s[r], kappa * mu / s[r]);
  }
  return {std::move(yv), std::move(zv)};
}

/// feasibility_restoration.hpp:49-100.
inline std::pair<Vec, Vec> compute_p_n(const Vec& c, double rho, double mu) {
  Vec p(c.size()), n(c.size());
  for (size_t r = 0; r < c.size(); ++r) {
    double a = rho;
    double b = rho * c[r] - mu;
    double cc = -mu * c[r] / 2.0;
    n[r] = (-b + std::sqrt(b * b - 4.0 * a * cc)) / (2.0 * a);
    p[r] = c[r] + n[r];
  }
  return {std::move(p), std::move(n)};
}

/// feasibility_restoration.hpp:346-628 (IPM variant).
inline ExitStatus feasibility_restoration(
    const IpmCallbacks& matrices, std::vector<IterationCallback>& callbacks,
    const Options& options, Vec& x, Vec& s, Vec& y, Vec& z, double mu,
    int& iterations, Trace* trace, const LinearSolverConfig& lin) {
  const int num_vars = matrices.num_decision_variables;
  const int num_eq = matrices.num_equality_constraints;
  const int num_ineq = matrices.num_inequality_constraints;
  constexpr double rho = 1e3;

  const Vec c_e = matrices.c_e(x);
  const Vec c_i = matrices.c_i(x);
  const Vec cis0 = sub(c_i, s);
  double fr_mu = std::max({mu, norm_inf(c_e), norm_inf(cis0)});
  const double zeta = std::sqrt(fr_mu);

  const Vec x_r = x;
  auto [p_e_0, n_e_0] = compute_p_n(c_e, rho, fr_mu);
  auto [p_i_0, n_i_0] = compute_p_n(cis0, rho, fr_mu);

  Vec D_r(num_vars);
  for (int i = 0; i < num_vars; ++i) {
    D_r[i] = std::min(1.0 / (x[i] * x[i]), 1.0);
  }

  const int extra = 2 * num_eq + 2 * num_ineq;
  Vec fr_x;
  fr_x.reserve(num_vars + extra);
  for (const Vec* v : {&x, &p_e_0, &n_e_0, &p_i_0, &n_i_0}) {
    fr_x.insert(fr_x.end(), v->begin(), v->end());
  }
  Vec fr_s(s.size() + extra, 1.0);
  std::copy(s.begin(), s.end(), fr_s.begin());
  Vec fr_y(c_e.size(), 0.0);
  Vec fr_z;
  fr_z.reserve(c_i.size() + extra);
  for (const Vec* v : {&s, &p_e_0, &n_e_0, &p_i_0, &n_i_0}) {
    for (double e : *v) fr_z.push_back(fr_mu * (1.0 / e));
  }

  Vec fr_d_c_i = matrices.scaling.c_i;
  fr_d_c_i.resize(c_i.size() + extra, 1.0);
  ProblemScaling fr_scaling{1.0, matrices.scaling.c_e, fr_d_c_i};

  const int fr_n = num_vars + extra;
  auto head = [&](const Vec& v, int off, int len) {
    return Vec(v.begin() + off, v.begin() + off + len);
  };

  IpmCallbacks fr;
  fr.num_decision_variables = fr_n;
  fr.num_equality_constraints = num_eq;
  fr.num_inequality_constraints = num_ineq + extra;
  fr.scaling = fr_scaling;
  fr.f = [&](const Vec& x_p) {
    double sum = 0.0;
    for (int i = 0; i < extra; ++i) sum += x_p[num_vars + i];
    double quad = 0.0;
    for (int i = 0; i < num_vars; ++i) {
      double d = x_p[i] - x_r[i];
      quad += d * (D_r[i] * d);
    }
    return rho * sum + zeta / 2.0 * quad;
  };
  fr.g = [&](const Vec& x_p) {
    Vec g(fr_n, rho);
    for (int i = 0; i < num_vars; ++i) {
      g[i] = zeta * D_r[i] * (x_p[i] - x_r[i]);
    }
    return g;
  };
  fr.H = [&](const Vec& x_p, const Vec& y_p, const Vec& z_p) {
    Vec d(num_vars);
    for (int i = 0; i < num_vars; ++i) d[i] = zeta * D_r[i];
    Csc d2f = diag(d).resized(fr_n, fr_n);
    // The reference evaluates H_c here and then calls Eigen's
    // SparseMatrix::resize() on it (:485-486), which re-initialises the matrix
    // to zero, so the restoration Hessian it actually factors is ζD_r alone.
    // Restated as it behaves; the callback is still invoked for its side
    // effects on the autodiff leaves.
    (void)matrices.H_c(head(x_p, 0, num_vars), y_p, head(z_p, 0, num_ineq));
    return add(d2f, Csc{fr_n, fr_n});
  };
  fr.H_c = [&](const Vec&, const Vec&, const Vec&) { return Csc{fr_n, fr_n}; };
  fr.c_e = [&](const Vec& x_p) {
    Vec c = matrices.c_e(head(x_p, 0, num_vars));
    for (int i = 0; i < num_eq; ++i) {
      c[i] = c[i] - x_p[num_vars + i] + x_p[num_vars + num_eq + i];
    }
    return c;
  };
  fr.A_e = [&](const Vec& x_p) {
    Csc A = matrices.A_e(head(x_p, 0, num_vars));
    Csc out{num_eq, fr_n};
    out.colptr = A.colptr;
    out.rowidx = A.rowidx;
    out.val = A.val;
    out.colptr.resize(fr_n + 1, A.nnz());
    for (int j = 0; j < num_eq; ++j) {  // −I block
      out.rowidx.push_back(j);
      out.val.push_back(-1.0);
      out.colptr[num_vars + j + 1] = out.nnz();
    }
    for (int j = 0; j < num_eq; ++j) {  // +I block
      out.rowidx.push_back(j);
      out.val.push_back(1.0);
      out.colptr[num_vars + num_eq + j + 1] = out.nnz();
    }
    for (int c = num_vars + 2 * num_eq; c < fr_n; ++c) {
      out.colptr[c + 1] = out.nnz();
    }
    return out;
  };
  fr.c_i = [&](const Vec& x_p) {
    Vec c = matrices.c_i(head(x_p, 0, num_vars));
    Vec out(num_ineq + extra);
    const int off_p = num_vars + 2 * num_eq;
    for (int i = 0; i < num_ineq; ++i) {
      out[i] = c[i] - x_p[off_p + i] + x_p[off_p + num_ineq + i];
    }
    for (int i = 0; i < extra; ++i) out[num_ineq + i] = x_p[num_vars + i];
    return out;
  };
  fr.A_i = [&](const Vec& x_p) {
    Csc A = matrices.A_i(head(x_p, 0, num_vars));
    Csc out{num_ineq + extra, fr_n};
    out.colptr = A.colptr;
    out.rowidx = A.rowidx;
    out.val = A.val;
    out.colptr.resize(fr_n + 1, A.nnz());
    // p_e, n_e columns: identity rows num_ineq + j
    for (int j = 0; j < 2 * num_eq; ++j) {
      out.rowidx.push_back(num_ineq + j);
      out.val.push_back(1.0);
      out.colptr[num_vars + j + 1] = out.nnz();
    }
    // p_i columns: −I on top, +I on their own bound rows
    for (int j = 0; j < num_ineq; ++j) {
      out.rowidx.push_back(j);
      out.val.push_back(-1.0);
      out.rowidx.push_back(num_ineq + 2 * num_eq + j);
      out.val.push_back(1.0);
      out.colptr[num_vars + 2 * num_eq + j + 1] = out.nnz();
    }
    // n_i columns: +I on top, +I on their own bound rows
    for (int j = 0; j < num_ineq; ++j) {
      out.rowidx.push_back(j);
      out.val.push_back(1.0);
      out.rowidx.push_back(num_ineq + 2 * num_eq + num_ineq + j);
      out.val.push_back(1.0);
      out.colptr[num_vars + 2 * num_eq + num_ineq + j + 1] = out.nnz();
    }
    return out;
  };

  // The restoration problem has a different KKT pattern: never reuse the
  // outer permutation.
  LinearSolverConfig fr_lin;
  fr_lin.ordering = lin.ordering == SimplicialLDLT::Ordering::CUSTOM
                        ? SimplicialLDLT::Ordering::AMD
                        : lin.ordering;
  fr_lin.force_sparse = lin.force_sparse;

  ExitStatus status =
      interior_point(fr, callbacks, options, true, fr_x, fr_s, fr_y, fr_z,
                     fr_mu, iterations, trace, fr_lin);

  x = head(fr_x, 0, num_vars);
  s = head(fr_s, 0, static_cast<int>(s.size()));

  if (status == ExitStatus::CALLBACK_REQUESTED_STOP) {
    Vec g = matrices.g(x);
    Csc A_e = matrices.A_e(x);
    Csc A_i = matrices.A_i(x);
    auto [ye, ze] = lagrange_multiplier_estimate(g, A_e, A_i, s, mu);
    y = ye;
    z = ze;
    return ExitStatus::SUCCESS;
  } else if (status == ExitStatus::SUCCESS) {
    return ExitStatus::LOCALLY_INFEASIBLE;
  }
  return ExitStatus::FEASIBILITY_RESTORATION_FAILED;
}

/// interior_point.hpp:123-866.
inline ExitStatus interior_point(const IpmCallbacks& matrices,
                                 std::vector<IterationCallback>& callbacks,
                                 const Options& options,
                                 bool in_feasibility_restoration, Vec& x,
                                 Vec& s, Vec& y, Vec& z, double& mu,
                                 int& iterations, Trace* trace,
                                 const LinearSolverConfig& lin,
                                 SolverKind kind) {
  struct Step {
    Vec p_x, p_s, p_y, p_z;
  };
  const auto solve_start = std::chrono::steady_clock::now();
  const int n = matrices.num_decision_variables;
  const int me = matrices.num_equality_constraints;

  double f = matrices.f(x);
  Vec g = matrices.g(x);
  Csc H = matrices.H(x, y, z);
  Vec c_e = matrices.c_e(x);
  Csc A_e = matrices.A_e(x);
  Vec c_i = matrices.c_i(x);
  Csc A_i = matrices.A_i(x);

  Vec trial_x, trial_s, trial_y, trial_z, trial_c_e, trial_c_i;
  double trial_f = 0.0;

  if (me > n) return ExitStatus::TOO_FEW_DOFS;  // :274-280

  if (!std::isfinite(f) || !all_finite(g) || !H.all_finite() ||
      !all_finite(c_e) || !A_e.all_finite() || !all_finite(c_i) ||
      !A_i.all_finite()) {
    return ExitStatus::NONFINITE_INITIAL_GUESS;  // :283-286
  }

  const double mu_min = matrices.scaling.f * options.tolerance / 10.0;
  constexpr double tau_min = 0.99;
  double tau = tau_min;

  Filter filter{norm1(c_e) + norm1(sub(c_i, s))};

  auto update_barrier_parameter_and_reset_filter = [&] {  // :308-333
    constexpr double kappa_mu = 0.2;
    constexpr double theta_mu = 1.5;
    mu = std::max(mu_min, std::min(kappa_mu * mu, std::pow(mu, theta_mu)));
    tau = std::max(tau_min, 1.0 - mu);
    filter.reset();
  };

  const int lhs_rows = n + me;
  bool use_sparse;
  if (lin.force_sparse >= 0) {
    use_sparse = lin.force_sparse != 0;
  } else {  // :340-348
    Csc AtA = matmul(A_i.transpose(), A_i).lower();
    use_sparse = H.nnz() + AtA.nnz() + A_e.nnz() <
                 0.25 * double(lhs_rows) * double(lhs_rows);
  }
  RegularizedLDLT solver{use_sparse, n, me,
                         in_feasibility_restoration ? 0.0 : 1e-10};
  if (lin.ordering == SimplicialLDLT::Ordering::CUSTOM) {
    solver.sparse_solver().set_custom_permutation(lin.permutation);
  } else {
    solver.sparse_solver().set_ordering(lin.ordering);
  }

  constexpr double alpha_reduction_factor = 0.5;
  constexpr double alpha_min = 1e-7;
  int full_step_rejected_counter = 0;

  double E_0 = unscaled_kkt_error(KKTErrorType::INF_NORM_SCALED,
                                  matrices.scaling, g, A_e, c_e, A_i, c_i, s, y,
                                  z, 0.0);

  while (E_0 > options.tolerance) {
    int it_solves = 0, it_trials = 0;
    const int fact_before = solver.factorizations;

    if (is_equality_locally_infeasible(A_e, c_e)) {
      return ExitStatus::LOCALLY_INFEASIBLE;
    }
    if (is_inequality_locally_infeasible(A_i, c_i)) {
      return ExitStatus::LOCALLY_INFEASIBLE;
    }
    if (norm_inf(x) > 1e10 || !all_finite(x) || norm_inf(s) > 1e10 ||
        !all_finite(s)) {
      return ExitStatus::DIVERGING_ITERATES;
    }

    for (const auto& cb : callbacks) {
      if (cb({iterations, x, s, y, z, g, H, A_e, A_i})) {
        return ExitStatus::CALLBACK_REQUESTED_STOP;
      }
    }

    // Σ = S⁻¹Z; lhs = [H + tril(AᵢᵀΣAᵢ); A_e] lower-only (:426-440)
    Vec sigma(s.size()), s_inv(s.size());
    for (size_t i = 0; i < s.size(); ++i) {
      s_inv[i] = 1.0 / s[i];
      sigma[i] = s_inv[i] * z[i];
    }
    Csc top_left =
        add(H, matmul(A_i.transpose().scale_cols(sigma), A_i).lower());
    Csc lhs = vstack(top_left, A_e).resized(lhs_rows, lhs_rows);

    // rhs = −[g − Aₑᵀy − Aᵢᵀ(−Σcᵢ + μS⁻¹e + z); cₑ] (:444-448)
    Vec rhs(lhs_rows);
    {
      Vec t(s.size());
      for (size_t i = 0; i < s.size(); ++i) {
        t[i] = -sigma[i] * c_i[i] + mu * s_inv[i] + z[i];
      }
      Vec aty = A_e.mul_t(y);
      Vec att = A_i.mul_t(t);
      for (int i = 0; i < n; ++i) rhs[i] = -g[i] + aty[i] + att[i];
      for (int i = 0; i < me; ++i) rhs[n + i] = -c_e[i];
    }

    Step step;
    double alpha_max = 1.0, alpha = 1.0, alpha_z = 1.0;
    bool call_feasibility_restoration = false;

    if (!solver.compute(lhs) && kind != SolverKind::NEWTON) {
      return ExitStatus::FACTORIZATION_FAILED;  // newton.hpp:183 ignores it
    }

    auto compute_step = [&](Step& st, const Vec& c_i_minus_s) {  // :470-481
      ++it_solves;
      Vec p = solver.solve(rhs);
      st.p_x.assign(p.begin(), p.begin() + n);
      st.p_y.resize(me);
      for (int i = 0; i < me; ++i) st.p_y[i] = -p[n + i];
      Vec aip = A_i.mul(st.p_x);
      st.p_s.resize(s.size());
      st.p_z.resize(s.size());
      for (size_t i = 0; i < s.size(); ++i) {
        st.p_s[i] = c_i_minus_s[i] + aip[i];
        st.p_z[i] = mu * s_inv[i] - z[i] - sigma[i] * st.p_s[i];
      }
    };
    compute_step(step, sub(c_i, s));

    alpha_max = fraction_to_the_boundary_rule(s, step.p_s, tau);
    alpha = alpha_max;
    if (alpha < alpha_min) call_feasibility_restoration = true;
    alpha_z = fraction_to_the_boundary_rule(z, step.p_z, tau);

    const FilterEntry current_entry{f, s, c_e, c_i, mu};
    const double D_phi = dot(g, step.p_x) - mu * dot(s_inv, step.p_s);

    while (true) {  // :512-717
      ++it_trials;
      trial_x = axpy(x, alpha, step.p_x);
      trial_c_i = matrices.c_i(trial_x);
      bool all_pos = std::all_of(c_i.begin(), c_i.end(),
                                 [](double v) { return v > 0.0; });
      if (options.feasible_ipm && all_pos) {
        trial_s = trial_c_i;
      } else {
        trial_s = axpy(s, alpha, step.p_s);
      }
      trial_y = axpy(y, kind == SolverKind::IPM ? alpha_z : alpha, step.p_y);
      trial_z = axpy(z, alpha_z, step.p_z);

      trial_f = matrices.f(trial_x);
      trial_c_e = matrices.c_e(trial_x);

      if (!std::isfinite(trial_f) || !all_finite(trial_c_e) ||
          !all_finite(trial_c_i)) {
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

      FilterEntry trial_entry{trial_f, trial_s, trial_c_e, trial_c_i, mu};
      if (filter.try_add(current_entry, trial_entry, D_phi, alpha)) break;

      double prev_violation = norm1(c_e) + norm1(sub(c_i, s));
      double next_violation =
          norm1(trial_c_e) + norm1(sub(trial_c_i, trial_s));

      // Second-order corrections (:561-664)
      if (kind != SolverKind::NEWTON && alpha == alpha_max &&
          next_violation >= prev_violation) {
        Step soc_step = step;
        double alpha_soc = alpha;
        double alpha_z_soc = alpha_z;
        Vec c_e_soc = c_e;
        Vec c_i_minus_s_soc = sub(c_i, s);
        double soc_violation = next_violation;
        bool step_acceptable = false;
        for (int soc_it = 0; soc_it < 5 && !step_acceptable; ++soc_it) {
          for (size_t i = 0; i < c_e_soc.size(); ++i) {
            c_e_soc[i] = alpha_soc * c_e_soc[i] + trial_c_e[i];
          }
          for (size_t i = 0; i < c_i_minus_s_soc.size(); ++i) {
            c_i_minus_s_soc[i] =
                alpha_soc * c_i_minus_s_soc[i] + trial_c_i[i] - trial_s[i];
          }
          {
            Vec t(s.size());
            for (size_t i = 0; i < s.size(); ++i) {
              t[i] = mu * s_inv[i] - sigma[i] * c_i_minus_s_soc[i];
            }
            Vec aty = A_e.mul_t(y);
            Vec att = A_i.mul_t(t);
            for (int i = 0; i < n; ++i) rhs[i] = -g[i] + aty[i] + att[i];
            for (int i = 0; i < me; ++i) rhs[n + i] = -c_e_soc[i];
          }
          compute_step(soc_step, c_i_minus_s_soc);
          alpha_soc = fraction_to_the_boundary_rule(s, soc_step.p_s, tau);
          alpha_z_soc = fraction_to_the_boundary_rule(z, soc_step.p_z, tau);

          trial_x = axpy(x, alpha_soc, soc_step.p_x);
          trial_s = axpy(s, alpha_soc, soc_step.p_s);
          trial_y = axpy(y, kind == SolverKind::IPM ? alpha_z_soc : alpha_soc,
                         soc_step.p_y);
          trial_z = axpy(z, alpha_z_soc, soc_step.p_z);

          ++it_trials;
          trial_f = matrices.f(trial_x);
          trial_c_e = matrices.c_e(trial_x);
          trial_c_i = matrices.c_i(trial_x);

          FilterEntry soc_entry{trial_f, trial_s, trial_c_e, trial_c_i, mu};
          if (filter.try_add(current_entry, soc_entry, D_phi, alpha)) {
            step = soc_step;
            alpha = alpha_soc;
            alpha_z = alpha_z_soc;
            step_acceptable = true;
            break;
          }
          constexpr double kappa_soc = 0.99;
          next_violation = norm1(trial_c_e) + norm1(sub(trial_c_i, trial_s));
          if (next_violation > kappa_soc * soc_violation) break;
          soc_violation = next_violation;
        }
        if (step_acceptable) break;
      }

      if (alpha == alpha_max) ++full_step_rejected_counter;

      if (kind != SolverKind::NEWTON && full_step_rejected_counter >= 4 &&
          filter.max_constraint_violation >
              current_entry.constraint_violation / 10.0 &&
          filter.last_rejection_due_to_filter()) {
        filter.max_constraint_violation *= 0.1;
        filter.reset();
        continue;
      }

      alpha *= alpha_reduction_factor;

      if (alpha < alpha_min) {  // :691-716
        double current_kkt_error = kkt_error(KKTErrorType::ONE_NORM, g, A_e,
                                             c_e, A_i, c_i, s, y, z, mu);
        trial_x = axpy(x, alpha_max, step.p_x);
        trial_s = axpy(s, alpha_max, step.p_s);
        trial_y = axpy(y, kind == SolverKind::IPM ? alpha_z : alpha_max,
                       step.p_y);
        trial_z = axpy(z, alpha_z, step.p_z);
        trial_f = matrices.f(trial_x);
        trial_c_e = matrices.c_e(trial_x);
        trial_c_i = matrices.c_i(trial_x);
        double next_kkt_error = kkt_error(
            KKTErrorType::ONE_NORM, matrices.g(trial_x), matrices.A_e(trial_x),
            trial_c_e, matrices.A_i(trial_x), trial_c_i, trial_s, trial_y,
            trial_z, mu);
        if (next_kkt_error <= 0.999 * current_kkt_error) break;
        if (kind == SolverKind::NEWTON) {
          return ExitStatus::LINE_SEARCH_FAILED;  // newton.hpp:245
        }
        call_feasibility_restoration = true;
        break;
      }
    }

    if (call_feasibility_restoration) {  // :721-771
      if (in_feasibility_restoration) {
        return ExitStatus::FEASIBILITY_RESTORATION_FAILED;
      }
      FilterEntry initial_entry{matrices.f(x), s, c_e, c_i, mu};
      std::vector<IterationCallback> fr_callbacks = callbacks;
      fr_callbacks.emplace_back([&](const IterationInfo& info) {
        Vec tx(info.x.begin(), info.x.begin() + n);
        Vec ts(info.s.begin(),
               info.s.begin() + matrices.num_inequality_constraints);
        Vec tce = matrices.c_e(tx);
        Vec tci = matrices.c_i(tx);
        FilterEntry te{matrices.f(tx), ts, tce, tci, mu};
        const double D_phi_r =
            dot(g, sub(tx, x)) - mu * dot(s_inv, sub(ts, s));
        return te.constraint_violation <
                   0.9 * initial_entry.constraint_violation &&
               filter.try_add(initial_entry, te, D_phi_r, alpha);
      });
      // the SQP variant enters with μ = tolerance/10
      // (feasibility_restoration.hpp:121)
      ExitStatus status = feasibility_restoration(
          matrices, fr_callbacks, options, x, s, y, z,
          kind == SolverKind::IPM ? mu : options.tolerance / 10.0, iterations,
          trace, lin);
      if (status != ExitStatus::SUCCESS) return status;
      f = matrices.f(x);
      c_e = matrices.c_e(x);
      c_i = matrices.c_i(x);
    } else {
      if (alpha == alpha_max) full_step_rejected_counter = 0;
      x = trial_x;
      s = trial_s;
      y = trial_y;
      z = trial_z;
      for (size_t r = 0; r < z.size(); ++r) {  // :797-801
        constexpr double kappa = 1e10;
        z[r] = std::clamp(z[r], 1.0 / kappa * mu / s[r], kappa * mu / s[r]);
      }
      f = trial_f;
      c_e = trial_c_e;
      c_i = trial_c_i;
    }

    // Re-linearise (:809-812)
    A_e = matrices.A_e(x);
    A_i = matrices.A_i(x);
    g = matrices.g(x);
    H = matrices.H(x, y, z);

    E_0 = unscaled_kkt_error(KKTErrorType::INF_NORM_SCALED, matrices.scaling,
                             g, A_e, c_e, A_i, c_i, s, y, z, 0.0);

    if (kind == SolverKind::IPM && E_0 > options.tolerance) {  // :819-832
      constexpr double kappa_eps = 10.0;
      double E_mu = kkt_error(KKTErrorType::INF_NORM_SCALED, g, A_e, c_e, A_i,
                              c_i, s, y, z, mu);
      while (mu > mu_min && E_mu <= kappa_eps * mu) {
        update_barrier_parameter_and_reset_filter();
        E_mu = kkt_error(KKTErrorType::INF_NORM_SCALED, g, A_e, c_e, A_i, c_i,
                         s, y, z, mu);
      }
    }

    if (trace != nullptr) {
      TraceRow row;
      row.iteration = iterations;
      row.type = in_feasibility_restoration ? 1 : 0;
      row.error = E_0;
      row.cost = f;
      row.infeasibility = norm1(c_e) + norm1(sub(c_i, s));
      row.complementarity = dot(s, z);
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
                      std::chrono::steady_clock::now() - solve_start)
                      .count();
      if (trace->keep_iterates) {
        row.x = x;
        row.s = s;
        row.y = y;
        row.z = z;
      }
      trace->total_factorizations += row.factorizations;
      trace->total_solves += it_solves;
      trace->total_trials += it_trials;
      trace->rows.push_back(std::move(row));
    }

    ++iterations;
    if (iterations >= options.max_iterations) {
      return ExitStatus::MAX_ITERATIONS_EXCEEDED;
    }
    if (std::chrono::duration<double>(std::chrono::steady_clock::now() -
                                      solve_start)
            .count() > options.timeout) {
      return ExitStatus::TIMEOUT;
    }
  }
  return ExitStatus::SUCCESS;
}

/// interior_point.hpp:63-87 — default s = 1, y = 0, z = 1, μ = 0.1·d_f.
inline ExitStatus interior_point(const IpmCallbacks& matrices,
                                 std::vector<IterationCallback>& callbacks,
                                 const Options& options, Vec& x,
                                 Trace* trace = nullptr,
                                 const LinearSolverConfig& lin = {},
                                 Vec* s_out = nullptr, Vec* y_out = nullptr,
                                 Vec* z_out = nullptr,
                                 SolverKind kind = SolverKind::IPM) {
  Vec s(matrices.num_inequality_constraints, 1.0);
  Vec y(matrices.num_equality_constraints, 0.0);
  Vec z(matrices.num_inequality_constraints, 1.0);
  double mu = 0.1 * matrices.scaling.f;
  int iterations = 0;
  ExitStatus st = interior_point(matrices, callbacks, options, false, x, s, y,
                                 z, mu, iterations, trace, lin, kind);
  if (s_out) *s_out = s;
  if (y_out) *y_out = y;
  if (z_out) *z_out = z;
  return st;
}

}  // namespace orc
