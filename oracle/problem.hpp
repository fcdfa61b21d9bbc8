// ORACLE — TEST INFRASTRUCTURE ONLY (see expr.hpp header).
//
// Restatement of include/sleipnir/optimization/problem.hpp: decision_variable
// (:78-101, row-major append), symmetric_decision_variable (:118-140),
// minimize/maximize (:151-190), subject_to (:196-234), and the interior-point
// branch of solve() (:281-313,512-679) including bound-conflict detection
// (solver/util/bounds.hpp:54-179), problem scaling (:615-616) and the eight
// matrix callbacks (:618-660). Problems without inequality constraints are
// dispatched to SQP/Newton by the reference (:335,403); those solvers are out
// of scope here, so solve() always takes the IPM branch.
#pragma once

#include <limits>
#include <memory>
#include <optional>
#include <utility>
#include <vector>

#include "autodiff.hpp"
#include "ipm.hpp"
#include "var.hpp"

namespace orc {

/// Everything Problem::solve builds before calling interior_point.
template <class B>
struct IpmSetup {
  Mat<B> x_ad, c_e_ad, c_i_ad, y_ad, z_ad;
  Var<B> f;
  std::unique_ptr<Gradient<B>> g;
  std::unique_ptr<Hessian<B>> H_f, H_c;
  std::unique_ptr<Jacobian<B>> A_e, A_i;
  ProblemScaling scaling;
  IpmCallbacks callbacks;
  std::vector<std::pair<int, int>> conflicting_bounds;
};

template <class B>
class Problem {
 public:
  using V = Var<B>;
  using M = Mat<B>;

  V decision_variable() {
    m_decision_variables.emplace_back();
    return m_decision_variables.back();
  }
  M decision_variable(int rows, int cols = 1) {
    M vars{typename M::empty_t{}, rows, cols};
    for (int r = 0; r < rows; ++r) {
      for (int c = 0; c < cols; ++c) {
        m_decision_variables.emplace_back();
        vars(r, c) = m_decision_variables.back();
      }
    }
    return vars;
  }
  M symmetric_decision_variable(int rows) {
    M vars{typename M::empty_t{}, rows, rows};
    for (int r = 0; r < rows; ++r) {
      for (int c = 0; c <= r; ++c) {
        m_decision_variables.emplace_back();
        vars(r, c) = m_decision_variables.back();
        vars(c, r) = m_decision_variables.back();
      }
    }
    return vars;
  }
  void minimize(const V& cost) { m_f = cost; }
  void maximize(const V& objective) { m_f = -objective; }
  void subject_to_eq(const std::vector<V>& c) {
    m_eq.insert(m_eq.end(), c.begin(), c.end());
  }
  void subject_to_ineq(const std::vector<V>& c) {
    m_ineq.insert(m_ineq.end(), c.begin(), c.end());
  }
  void add_callback(IterationCallback cb) {
    m_callbacks.push_back(std::move(cb));
  }
  void clear_callbacks() { m_callbacks.clear(); }

  int num_decision_variables() const {
    return static_cast<int>(m_decision_variables.size());
  }
  int num_equality_constraints() const { return static_cast<int>(m_eq.size()); }
  int num_inequality_constraints() const {
    return static_cast<int>(m_ineq.size());
  }
  std::vector<V>& decision_variables() { return m_decision_variables; }
  std::vector<V>& equality_constraints() { return m_eq; }
  std::vector<V>& inequality_constraints() { return m_ineq; }
  const std::optional<V>& cost() const { return m_f; }

  int cost_function_type() const { return m_f ? m_f->type() : 0; }
  int equality_constraint_type() const { return max_type(m_eq); }
  int inequality_constraint_type() const { return max_type(m_ineq); }

  Vec initial_guess() {
    Vec x(m_decision_variables.size());
    for (size_t i = 0; i < x.size(); ++i) {
      x[i] = m_decision_variables[i].value();
    }
    return x;
  }

  /// problem.hpp:517-660.
  std::unique_ptr<IpmSetup<B>> make_ipm_setup(const Vec& x0) {
    auto st = std::make_unique<IpmSetup<B>>();
    IpmSetup<B>& S = *st;
    S.x_ad = M{m_decision_variables};
    S.f = m_f.value_or(V{0.0});
    const int n = num_decision_variables();
    const int me = num_equality_constraints();
    const int mi = num_inequality_constraints();
    S.c_e_ad = M{m_eq};
    S.c_i_ad = M{m_ineq};
    S.y_ad = M(me, 1);
    S.z_ad = M(mi, 1);
    if (me == 0) S.c_e_ad = M{typename M::empty_t{}, 0, 1};
    if (mi == 0) S.c_i_ad = M{typename M::empty_t{}, 0, 1};

    S.g = std::make_unique<Gradient<B>>(S.f, S.x_ad);
    S.H_f = std::make_unique<Hessian<B>>(S.f, S.x_ad);
    // −yᵀcₑ − zᵀcᵢ as 1×1 matrix products (left-deep chains), :547-548.
    M lagr;
    {
      M ye = (-S.y_ad.T()) * S.c_e_ad;
      M zi = S.z_ad.T() * S.c_i_ad;
      lagr = ye - zi;
    }
    S.H_c = std::make_unique<Hessian<B>>(lagr(0, 0), S.x_ad);
    S.A_e = std::make_unique<Jacobian<B>>(S.c_e_ad, S.x_ad);
    S.A_i = std::make_unique<Jacobian<B>>(S.c_i_ad, S.x_ad);

    S.conflicting_bounds = detect_conflicting_bounds(S.A_i->value());

    S.x_ad.set_value(x0);
    S.scaling = ProblemScaling{S.g->value(), S.A_e->value(), S.A_i->value()};

    IpmSetup<B>* p = st.get();
    auto scaled = [](const Vec& d, const Vec& v) {
      Vec r(v.size());
      for (size_t i = 0; i < v.size(); ++i) r[i] = d[i] * v[i];
      return r;
    };
    auto& cb = S.callbacks;
    cb.num_decision_variables = n;
    cb.num_equality_constraints = me;
    cb.num_inequality_constraints = mi;
    cb.scaling = S.scaling;
    cb.f = [p](const Vec& x) {
      p->x_ad.set_value(x);
      return p->scaling.f * p->f.value();
    };
    cb.g = [p](const Vec& x) {
      p->x_ad.set_value(x);
      Vec g = p->g->value();
      for (auto& v : g) v = p->scaling.f * v;
      return g;
    };
    cb.H = [p, scaled](const Vec& x, const Vec& y, const Vec& z) {
      p->x_ad.set_value(x);
      p->y_ad.set_value(scaled(p->scaling.c_e, y));
      p->z_ad.set_value(scaled(p->scaling.c_i, z));
      return add(p->H_f->value().scaled(p->scaling.f), p->H_c->value());
    };
    cb.H_c = [p, scaled](const Vec& x, const Vec& y, const Vec& z) {
      p->x_ad.set_value(x);
      p->y_ad.set_value(scaled(p->scaling.c_e, y));
      p->z_ad.set_value(scaled(p->scaling.c_i, z));
      return p->H_c->value();
    };
    cb.c_e = [p, scaled](const Vec& x) {
      p->x_ad.set_value(x);
      return scaled(p->scaling.c_e, p->c_e_ad.value());
    };
    cb.A_e = [p](const Vec& x) {
      p->x_ad.set_value(x);
      return p->A_e->value().scale_rows(p->scaling.c_e);
    };
    cb.c_i = [p, scaled](const Vec& x) {
      p->x_ad.set_value(x);
      return scaled(p->scaling.c_i, p->c_i_ad.value());
    };
    cb.A_i = [p](const Vec& x) {
      p->x_ad.set_value(x);
      return p->A_i->value().scale_rows(p->scaling.c_i);
    };
    return st;
  }

  ExitStatus solve(const Options& options = {}, Trace* trace = nullptr,
                   const LinearSolverConfig& lin = {}, Vec* s_out = nullptr,
                   Vec* y_out = nullptr, Vec* z_out = nullptr) {
    Vec x = initial_guess();
    constexpr int CONSTANT = 1;
    if (cost_function_type() <= CONSTANT &&
        equality_constraint_type() <= CONSTANT &&
        inequality_constraint_type() <= CONSTANT) {
      return ExitStatus::SUCCESS;
    }
    auto setup = make_ipm_setup(x);
    if (!setup->conflicting_bounds.empty()) {
      return ExitStatus::GLOBALLY_INFEASIBLE;
    }
    std::vector<IterationCallback> cbs = m_callbacks;
    // problem.hpp:335 (Newton), :403 (SQP), :512 (IPM). The three drivers
    // share one setup here: with no inequality (equality) constraints the
    // z (y) blocks are empty and H_c's missing terms are exact zeros.
    SolverKind kind = SolverKind::IPM;
    if (m_eq.empty() && m_ineq.empty()) {
      kind = SolverKind::NEWTON;
    } else if (m_ineq.empty()) {
      kind = SolverKind::SQP;
    }
    ExitStatus status = interior_point(setup->callbacks, cbs, options, x, trace,
                                       lin, s_out, y_out, z_out, kind);
    M{m_decision_variables}.set_value(x);
    return status;
  }

 private:
  static int max_type(const std::vector<V>& v) {
    int t = 0;
    for (const auto& e : v) t = std::max(t, e.type());
    return t;
  }

  /// solver/util/bounds.hpp:54-179, reduced to the conflict list that
  /// Problem::solve consumes (:597-606).
  std::vector<std::pair<int, int>> detect_conflicting_bounds(const Csc& A_i) {
    constexpr int NO_BOUND = -1;
    constexpr int LINEAR = 2;
    const size_t n = m_decision_variables.size();
    std::vector<std::pair<int, int>> idx(n, {NO_BOUND, NO_BOUND});
    std::vector<std::pair<double, double>> bnd(
        n, {-std::numeric_limits<double>::infinity(),
            std::numeric_limits<double>::infinity()});
    std::vector<std::pair<int, int>> conflicts;
    Csc rows = A_i.transpose();  // column r of `rows` = row r of A_i
    for (int ci = 0; ci < static_cast<int>(m_ineq.size()); ++ci) {
      if (m_ineq[ci].type() != LINEAR) continue;
      int nz = rows.colptr[ci + 1] - rows.colptr[ci];
      if (nz != 1) continue;
      double coeff = rows.val[rows.colptr[ci]];
      int var = rows.rowidx[rows.colptr[ci]];
      double var_value = m_decision_variables[var].value();
      double constant;
      if (var_value != 0.0) {
        m_decision_variables[var].set_value(0.0);
        constant = m_ineq[ci].value();
        m_decision_variables[var].set_value(var_value);
      } else {
        constant = m_ineq[ci].value();
      }
      auto& [lo, hi] = bnd[var];
      auto& [lo_i, hi_i] = idx[var];
      double detected = -constant / coeff;
      if (coeff < 0.0 && detected < hi) {
        hi = detected;
        hi_i = ci;
      } else if (coeff > 0.0 && detected > lo) {
        lo = detected;
        lo_i = ci;
      }
      if (lo > hi) conflicts.emplace_back(lo_i, hi_i);
    }
    return conflicts;
  }

  std::vector<V> m_decision_variables;
  std::optional<V> m_f;
  std::vector<V> m_eq, m_ineq;
  std::vector<IterationCallback> m_callbacks;
};

}  // namespace orc
