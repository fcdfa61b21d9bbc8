// ORACLE — TEST INFRASTRUCTURE ONLY (see expr.hpp header).
//
// extern "C" surface over the CPU oracle, loaded with ctypes by tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
// Built twice from this one file:
//   oracle/liboracle.so           restated expression core (expr.hpp)
//   oracle/_ref/liboracle_ref.so  -DORC_BACKEND_REF: the reference's own
//                                 expression.hpp / expression_graph.hpp
#include <chrono>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#ifdef ORC_BACKEND_REF
#include "backend_ref.hpp"
#else
#include "expr.hpp"
#endif
#include "problems.hpp"

namespace {

#ifdef ORC_BACKEND_REF
using Bk = orc::RefBackend;
#else
using Bk = orc::OwnBackend;
#endif

struct Handle {
  std::unique_ptr<orc::Problem<Bk>> problem;
  std::unique_ptr<orc::IpmSetup<Bk>> setup;
  orc::Trace trace;
  orc::Vec x, s, y, z;
  orc::Csc last;  // last matrix evaluated through orc_eval_matrix
  double solve_seconds = 0.0;
  int iterations = 0;
};

Handle* H(void* h) { return static_cast<Handle*>(h); }

int g_live_handles = 0;

orc::Vec vec(const double* p, int n) { return orc::Vec(p, p + n); }

}  // namespace

extern "C" {

const char* orc_backend_name() { return Bk::name(); }

void* orc_problem_create(const char* name, int N, double p0, double p1) {
  try {
    auto h = std::make_unique<Handle>();
    h->problem = orc::make_problem<Bk>(name, N, p0, p1);
    ++g_live_handles;
    return h.release();
  } catch (...) {
    return nullptr;
  }
}

void orc_problem_destroy(void* h) {
  delete H(h);
  --g_live_handles;
#ifndef ORC_BACKEND_REF
  // The restated core allocates nodes from an arena; release it once no
  // problem references it any more.
  if (g_live_handles == 0) orc::arena().reset();
#endif
}

void orc_problem_dims(void* h, int* n, int* me, int* mi) {
  *n = H(h)->problem->num_decision_variables();
  *me = H(h)->problem->num_equality_constraints();
  *mi = H(h)->problem->num_inequality_constraints();
}

void orc_problem_types(void* h, int* f, int* ce, int* ci) {
  *f = H(h)->problem->cost_function_type();
  *ce = H(h)->problem->equality_constraint_type();
  *ci = H(h)->problem->inequality_constraint_type();
}

void orc_problem_initial_guess(void* h, double* x) {
  orc::Vec v = H(h)->problem->initial_guess();
  std::memcpy(x, v.data(), v.size() * sizeof(double));
}

void orc_problem_set_guess(void* h, const double* x) {
  auto& vars = H(h)->problem->decision_variables();
  for (size_t i = 0; i < vars.size(); ++i) vars[i].set_value(x[i]);
}

/// ordering: 0 AMD (Eigen default), 1 natural, 2 custom (perm[k] = original
/// index eliminated k-th). Returns the ExitStatus value.
int orc_problem_solve(void* h, double tolerance, int max_iterations,
                      int feasible_ipm, int ordering, const int* perm,
                      int force_sparse, int keep_iterates) {
  Handle* hd = H(h);
  orc::Options opt;
  opt.tolerance = tolerance;
  opt.max_iterations = max_iterations;
  opt.feasible_ipm = feasible_ipm != 0;
  orc::LinearSolverConfig lin;
  lin.force_sparse = force_sparse;
  if (ordering == 1) {
    lin.ordering = orc::SimplicialLDLT::Ordering::NATURAL;
  } else if (ordering == 2) {
    lin.ordering = orc::SimplicialLDLT::Ordering::CUSTOM;
    int dim = hd->problem->num_decision_variables() +
              hd->problem->num_equality_constraints();
    lin.permutation.assign(perm, perm + dim);
  }
  hd->trace = orc::Trace{};
  hd->trace.keep_iterates = keep_iterates != 0;
  auto t0 = std::chrono::steady_clock::now();
  orc::ExitStatus st =
      hd->problem->solve(opt, &hd->trace, lin, &hd->s, &hd->y, &hd->z);
  hd->solve_seconds =
      std::chrono::duration<double>(std::chrono::steady_clock::now() - t0)
          .count();
  hd->x = hd->problem->initial_guess();  // solution was written back
  hd->iterations = static_cast<int>(hd->trace.rows.size());
  return static_cast<int>(st);
}

double orc_solve_seconds(void* h) { return H(h)->solve_seconds; }
int orc_trace_rows(void* h) { return static_cast<int>(H(h)->trace.rows.size()); }

/// scalars[16]: iteration, type, error, cost, infeasibility, complementarity,
/// mu, delta, gamma, alpha, alpha_max, alpha_z, factorizations, solves, trials.
void orc_trace_get(void* h, int row, double* scalars, double* x, double* s,
                   double* y, double* z) {
  const orc::TraceRow& r = H(h)->trace.rows[row];
  double sc[16] = {double(r.iteration), double(r.type), r.error, r.cost,
                   r.infeasibility, r.complementarity, r.mu, r.delta, r.gamma,
                   r.alpha, r.alpha_max, r.alpha_z, double(r.factorizations),
                   double(r.solves), double(r.trials), r.t_end};
  std::memcpy(scalars, sc, sizeof(sc));
  auto cp = [](double* dst, const orc::Vec& v) {
    if (dst && !v.empty()) std::memcpy(dst, v.data(), v.size() * sizeof(double));
  };
  cp(x, r.x);
  cp(s, r.s);
  cp(y, r.y);
  cp(z, r.z);
}

void orc_solution(void* h, double* x, double* s, double* y, double* z) {
  Handle* hd = H(h);
  auto cp = [](double* dst, const orc::Vec& v) {
    if (dst && !v.empty()) std::memcpy(dst, v.data(), v.size() * sizeof(double));
  };
  cp(x, hd->x);
  cp(s, hd->s);
  cp(y, hd->y);
  cp(z, hd->z);
}

// ---- callback-level evaluation (kernel parity) -----------------------------

/// Builds Gradient/Hessian/Jacobian objects and the scaling at the current
/// guess (problem.hpp:517-660). Returns the number of conflicting bounds.
int orc_eval_setup(void* h) {
  Handle* hd = H(h);
  hd->setup = hd->problem->make_ipm_setup(hd->problem->initial_guess());
  return static_cast<int>(hd->setup->conflicting_bounds.size());
}

void orc_eval_scaling(void* h, double* d_f, double* d_ce, double* d_ci) {
  const auto& sc = H(h)->setup->scaling;
  *d_f = sc.f;
  if (!sc.c_e.empty()) std::memcpy(d_ce, sc.c_e.data(), sc.c_e.size() * 8);
  if (!sc.c_i.empty()) std::memcpy(d_ci, sc.c_i.data(), sc.c_i.size() * 8);
}

double orc_eval_f(void* h, const double* x) {
  auto& cb = H(h)->setup->callbacks;
  return cb.f(vec(x, cb.num_decision_variables));
}

/// which: 0 g (n), 1 c_e (me), 2 c_i (mi)
void orc_eval_vector(void* h, int which, const double* x, double* out) {
  auto& cb = H(h)->setup->callbacks;
  orc::Vec xv = vec(x, cb.num_decision_variables);
  orc::Vec r = which == 0 ? cb.g(xv) : which == 1 ? cb.c_e(xv) : cb.c_i(xv);
  if (!r.empty()) std::memcpy(out, r.data(), r.size() * sizeof(double));
}

/// which: 0 A_e, 1 A_i, 2 H, 3 H_c. Result kept in the handle; returns nnz.
int orc_eval_matrix(void* h, int which, const double* x, const double* y,
                    const double* z) {
  Handle* hd = H(h);
  auto& cb = hd->setup->callbacks;
  orc::Vec xv = vec(x, cb.num_decision_variables);
  if (which == 0) {
    hd->last = cb.A_e(xv);
  } else if (which == 1) {
    hd->last = cb.A_i(xv);
  } else {
    orc::Vec yv = vec(y, cb.num_equality_constraints);
    orc::Vec zv = vec(z, cb.num_inequality_constraints);
    hd->last = which == 2 ? cb.H(xv, yv, zv) : cb.H_c(xv, yv, zv);
  }
  return hd->last.nnz();
}

void orc_last_matrix(void* h, int* rows, int* cols, int* colptr, int* rowidx,
                     double* val) {
  const orc::Csc& m = H(h)->last;
  *rows = m.rows;
  *cols = m.cols;
  std::memcpy(colptr, m.colptr.data(), m.colptr.size() * sizeof(int));
  if (m.nnz() > 0) {
    std::memcpy(rowidx, m.rowidx.data(), m.rowidx.size() * sizeof(int));
    std::memcpy(val, m.val.data(), m.val.size() * sizeof(double));
  }
}

/// Times `reps` re-linearisations + one trial evaluation each, the way one
/// Newton iteration calls them (interior_point.hpp:514,527-528,809-812).
/// Returns seconds per iteration-equivalent of graph walking.
double orc_time_graph_walk(void* h, const double* x, const double* y,
                           const double* z, int reps) {
  Handle* hd = H(h);
  auto& cb = hd->setup->callbacks;
  orc::Vec xv = vec(x, cb.num_decision_variables);
  orc::Vec yv = vec(y, cb.num_equality_constraints);
  orc::Vec zv = vec(z, cb.num_inequality_constraints);
  auto t0 = std::chrono::steady_clock::now();
  double sink = 0.0;
  for (int r = 0; r < reps; ++r) {
    sink += cb.c_i(xv)[0];
    sink += cb.f(xv);
    sink += cb.c_e(xv)[0];
    sink += cb.A_e(xv).nnz();
    sink += cb.A_i(xv).nnz();
    sink += cb.g(xv)[0];
    sink += cb.H(xv, yv, zv).nnz();
  }
  double dt =
      std::chrono::duration<double>(std::chrono::steady_clock::now() - t0)
          .count();
  return sink == -1.0 ? 0.0 : dt / reps;
}

// ---- linear algebra probes (ordering experiments, LDLᵀ parity) -------------

/// AMD permutation of a symmetric matrix given by its lower triangle.
void orc_amd(int n, const int* colptr, const int* rowidx, int* perm_out) {
  orc::Csc lower{n, n};
  lower.colptr.assign(colptr, colptr + n + 1);
  lower.rowidx.assign(rowidx, rowidx + colptr[n]);
  lower.val.assign(colptr[n], 0.0);
  std::vector<int> Ap, Ai;
  orc::symmetrize_pattern(lower, Ap, Ai);
  std::vector<int> p = orc::amd_order(n, Ap, Ai);
  std::memcpy(perm_out, p.data(), n * sizeof(int));
}

/// Factor + solve a symmetric system (lower triangle CSC) with a given
/// elimination order (perm == nullptr → AMD). Outputs D (permuted order) and
/// the solution; returns nnz(L), or −1 if a pivot was exactly zero.
int orc_ldlt(int n, const int* colptr, const int* rowidx, const double* val,
             const int* perm, const double* rhs, double* D_out, double* x_out,
             int* etree_height) {
  orc::Csc lower{n, n};
  lower.colptr.assign(colptr, colptr + n + 1);
  lower.rowidx.assign(rowidx, rowidx + colptr[n]);
  lower.val.assign(val, val + colptr[n]);
  orc::SimplicialLDLT f;
  if (perm) f.set_custom_permutation(std::vector<int>(perm, perm + n));
  f.analyze(lower);
  if (etree_height) {
    const auto& par = f.parent();
    std::vector<int> depth(n, 0);
    int best = 0;
    for (int k = 0; k < n; ++k) {
      if (par[k] >= 0) depth[par[k]] = std::max(depth[par[k]], depth[k] + 1);
      best = std::max(best, depth[k] + 1);
    }
    *etree_height = best;
  }
  bool ok = f.factorize(lower);
  if (D_out) std::memcpy(D_out, f.vectorD().data(), n * sizeof(double));
  if (!ok) return -1;
  if (rhs && x_out) {
    orc::Vec x = f.solve(orc::Vec(rhs, rhs + n));
    std::memcpy(x_out, x.data(), n * sizeof(double));
  }
  return f.nnz_l();
}

}  // extern "C"
