// libslpb_host.so — C API over the host side (slp::Problem DSL + IPM driver +
// the benchmark/parity problem builders) so that Python (tests/, bench.py,
// __graft_entry__.py) can drive it with ctypes. The host side is C++ because
// the reference's is; everything numerical below it goes through the C ABI of
// include/slpb.h into the CUDA library.
#include <algorithm>
#include <array>
#include <chrono>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "problems/problems.hpp"
#include "sleipnir/optimization/multistart.hpp"

namespace {

struct Handle {
  std::unique_ptr<slp::Problem<double>> problem;
  std::unique_ptr<slp::Problem<double>::Graphs> graphs;
  std::unique_ptr<slp::DeviceHandle> device;
  std::string error;
  std::vector<double> x;
  int last_status = 0;
  bool flush_l2 = false;
  int factor_arithmetic = -1;
  bool diagnostics = false, spy = false;
  double timeout_s = -1.0;  // < 0: Options default (no timeout)
  int rank = 0, world = 1;
  std::array<char, 128> nccl_id{};
  // what the recording callback saw, 8 doubles per call: iteration, n, ‖x‖∞,
  // s.size, y.size, z.size, nnz(H) + nnz(A_e) + nnz(A_i), g.size
  std::vector<double> callback_log;
  std::vector<double> callback_last_x;
};

Handle* H(void* h) { return static_cast<Handle*>(h); }

}  // namespace

extern "C" {

void* slpbh_problem_create(const char* name, int N, double p0, double p1) {
  try {
    auto h = std::make_unique<Handle>();
    h->problem = slpb_problems::make_problem(name, N, p0, p1);
    return h.release();
  } catch (...) {
    return nullptr;
  }
}

void slpbh_problem_destroy(void* h) { delete H(h); }

const char* slpbh_error(void* h) { return H(h)->error.c_str(); }

void slpbh_dims(void* h, int* n, int* me, int* mi) {
  *n = static_cast<int>(H(h)->problem->decision_variables().size());
  *me = static_cast<int>(H(h)->problem->equality_constraints().size());
  *mi = static_cast<int>(H(h)->problem->inequality_constraints().size());
}

void slpbh_types(void* h, int* f, int* ce, int* ci) {
  *f = static_cast<int>(H(h)->problem->cost_function_type());
  *ce = static_cast<int>(H(h)->problem->equality_constraint_type());
  *ci = static_cast<int>(H(h)->problem->inequality_constraint_type());
}

void slpbh_initial_guess(void* h, double* x) {
  auto& vars = H(h)->problem->decision_variables();
  for (size_t i = 0; i < vars.size(); ++i) x[i] = vars[i].value();
}

void slpbh_set_guess(void* h, const double* x) {
  auto& vars = H(h)->problem->decision_variables();
  for (size_t i = 0; i < vars.size(); ++i) vars[i].set_value(x[i]);
}

/// Problem::solve. Returns the ExitStatus value, or −100 when the device
/// library failed (slpbh_error has the text).
int slpbh_solve(void* h, double tolerance, int max_iterations, int feasible_ipm,
                int device, int ordering, const int* perm, int keep_iterates) {
  Handle* hd = H(h);
  slp::Options opt;
  opt.tolerance = tolerance;
  opt.max_iterations = max_iterations;
  opt.feasible_ipm = feasible_ipm != 0;
  if (hd->timeout_s >= 0.0) {
    opt.timeout = std::chrono::duration<double>(hd->timeout_s);
  }
  slp::DeviceOptions dopt;
  dopt.device = device;
  dopt.ordering = ordering;
  dopt.keep_iterates = keep_iterates != 0;
  dopt.flush_l2 = hd->flush_l2;
  dopt.factor_arithmetic = hd->factor_arithmetic;
  dopt.spy = hd->spy;
  opt.diagnostics = hd->diagnostics;
  dopt.rank = hd->rank;
  dopt.world = hd->world;
  dopt.nccl_unique_id = hd->nccl_id;
  if (perm) {
    const size_t dim = hd->problem->decision_variables().size() +
                       hd->problem->equality_constraints().size();
    dopt.permutation.assign(perm, perm + dim);
  }
  try {
    hd->last_status = static_cast<int>(hd->problem->solve(opt, dopt));
  } catch (const std::exception& e) {
    hd->error = e.what();
    return -100;
  }
  return hd->last_status;
}

/// slp::multistart (reference multistart.hpp:44-73) over `count` starts of a
/// named problem, each built and solved on its own thread through its own
/// device handle. Start i is make_problem(name, N, p0[i], p1[i]) with its
/// initial guess shifted by perturb·i·sin(1 + j) in component j (0 keeps the
/// builder's guess). best[3] = status, cost, winning index; best_x = its
/// decision variables (x_cap doubles at most); per_start[4·i..] = status, cost,
/// iterations, seconds inside solve(). Returns the wall time of the whole
/// multistart call in seconds, or −1 when a device call failed.
double slpbh_multistart(const char* name, int N, int count, const double* p0,
                        const double* p1, double perturb, int max_concurrency,
                        int max_iterations, int device, double* best,
                        double* best_x, int x_cap, double* per_start) {
  struct Start {
    int index;
    double p0, p1;
  };
  struct Vars {
    int index = 0;
    int iterations = 0;
    double seconds = 0.0;
    std::vector<double> x;
  };
  const std::string problem_name{name};
  std::vector<Start> starts(count);
  for (int i = 0; i < count; ++i) starts[i] = {i, p0[i], p1[i]};
  using Result = slp::MultistartResult<double, Vars>;
  const std::function<Result(const Start&)> solve_one =
      [&](const Start& st) -> Result {
    auto problem = slpb_problems::make_problem(problem_name, N, st.p0, st.p1);
    auto& vars = problem->decision_variables();
    if (perturb != 0.0) {
      for (size_t j = 0; j < vars.size(); ++j) {
        vars[j].set_value(vars[j].value() +
                          perturb * st.index * std::sin(1.0 + double(j)));
      }
    }
    slp::Options opt;
    opt.max_iterations = max_iterations;
    slp::DeviceOptions dopt;
    dopt.device = device;
    const auto t0 = std::chrono::steady_clock::now();
    const slp::ExitStatus status = problem->solve(opt, dopt);
    Vars out;
    out.index = st.index;
    out.seconds = std::chrono::duration<double>(
                      std::chrono::steady_clock::now() - t0)
                      .count();
    out.iterations = static_cast<int>(problem->last_trace().rows.size());
    out.x.resize(vars.size());
    for (size_t j = 0; j < vars.size(); ++j) out.x[j] = vars[j].value();
    auto cost = problem->cost();
    return {status, cost ? cost->value() : 0.0, std::move(out)};
  };
  // multistart() is written over DecisionVariables for both the guess and the
  // result; adapt by carrying the start inside Vars
  std::vector<Vars> guesses(count);
  for (int i = 0; i < count; ++i) guesses[i].index = i;
  const std::function<Result(const Vars&)> solve =
      [&](const Vars& g) { return solve_one(starts[g.index]); };
  std::vector<Result> all;
  const auto t0 = std::chrono::steady_clock::now();
  try {
    // SLPB_NO_GROUP=1: every start factors and solves on its own (round 1's
    // behaviour), for comparisons
    const int batch_device = std::getenv("SLPB_NO_GROUP") ? -1 : device;
    const Result win = slp::multistart<double, Vars>(
        solve, std::span<const Vars>{guesses}, max_concurrency, &all,
        batch_device);
    const double wall = std::chrono::duration<double>(
                            std::chrono::steady_clock::now() - t0)
                            .count();
    best[0] = static_cast<double>(static_cast<int>(win.status));
    best[1] = win.cost;
    best[2] = win.variables.index;
    for (size_t j = 0; j < win.variables.x.size() && int(j) < x_cap; ++j) {
      best_x[j] = win.variables.x[j];
    }
    for (int i = 0; i < count; ++i) {
      per_start[4 * i + 0] = static_cast<double>(static_cast<int>(all[i].status));
      per_start[4 * i + 1] = all[i].cost;
      per_start[4 * i + 2] = all[i].variables.iterations;
      per_start[4 * i + 3] = all[i].variables.seconds;
    }
    return wall;
  } catch (const std::exception&) {
    return -1.0;
  }
}

/// Benchmark hygiene: evict the device L2 before every iteration of the next
/// solves (excluded from the iteration timestamps).
void slpbh_set_flush_l2(void* h, int on) { H(h)->flush_l2 = on != 0; }
/// DeviceOptions::factor_arithmetic of the following solve() calls.
void slpbh_set_factor_arithmetic(void* h, int mode) {
  H(h)->factor_arithmetic = mode;
}
double slpbh_flush_seconds(void* h) {
  return H(h)->problem->last_trace().flush_seconds;
}

/// out[9]: build_graphs, flatten, device_create, upload+compile, scaling,
/// analyze, newton loop, write-back, teardown — host seconds of the last
/// solve().
void slpbh_phase_seconds(void* h, double* out) {
  const auto& p = H(h)->problem->last_phase_seconds();
  for (int i = 0; i < 9; ++i) out[i] = p[i];
}

/// Multi-GPU sharded solves: this process is `rank` of `world`, id = the
/// 128-byte ncclUniqueId of slpb_comm_unique_id distributed by the caller.
void slpbh_set_comm(void* h, int rank, int world, const void* id) {
  Handle* hd = H(h);
  hd->rank = rank;
  hd->world = world;
  if (id) std::memcpy(hd->nccl_id.data(), id, 128);
}

/// Options::diagnostics (iteration table + timing tables on stdout) and the
/// `spy` argument of Problem::solve (H.spy / A_e.spy / A_i.spy in the working
/// directory) for the next solves.
void slpbh_set_diagnostics(void* h, int diagnostics, int spy) {
  H(h)->diagnostics = diagnostics != 0;
  H(h)->spy = spy != 0;
}

/// Options::timeout for the next solves (seconds; negative = none).
void slpbh_set_timeout(void* h, double seconds) { H(h)->timeout_s = seconds; }

/// Registers an iteration callback (Problem::add_callback, or
/// add_persistent_callback) that records what IterationInfo hands it and asks
/// the solver to stop once `stop_at` iterations have been seen (stop_at < 0:
/// never).
void slpbh_add_recording_callback(void* h, int stop_at, int persistent) {
  Handle* hd = H(h);
  auto cb = [hd, stop_at](const slp::IterationInfo<double>& info) -> bool {
    double xinf = 0.0;
    for (int i = 0; i < static_cast<int>(info.x.size()); ++i) {
      xinf = std::max(xinf, std::abs(info.x[i]));
    }
    const double rec[8] = {double(info.iteration), double(info.x.size()), xinf,
                           double(info.s.size()), double(info.y.size()),
                           double(info.z.size()),
                           double(info.H.nonZeros() + info.A_e.nonZeros() +
                                  info.A_i.nonZeros()),
                           double(info.g.size())};
    hd->callback_log.insert(hd->callback_log.end(), rec, rec + 8);
    hd->callback_last_x.assign(info.x.data(), info.x.data() + info.x.size());
    return stop_at >= 0 && info.iteration >= stop_at;
  };
  if (persistent) {
    hd->problem->add_persistent_callback(cb);
  } else {
    hd->problem->add_callback(cb);
  }
}
void slpbh_clear_callbacks(void* h) { H(h)->problem->clear_callbacks(); }
int slpbh_callback_log(void* h, double* out, int max_records, double* last_x) {
  Handle* hd = H(h);
  const int n = static_cast<int>(hd->callback_log.size() / 8);
  for (int i = 0; i < std::min(n, max_records) * 8; ++i) {
    out[i] = hd->callback_log[i];
  }
  if (last_x && !hd->callback_last_x.empty()) {
    std::memcpy(last_x, hd->callback_last_x.data(),
                hd->callback_last_x.size() * 8);
  }
  return n;
}

int slpbh_trace_rows(void* h) {
  return static_cast<int>(H(h)->problem->last_trace().rows.size());
}

/// scalars[16]: same layout as the oracle's orc_trace_get.
void slpbh_trace_get(void* h, int row, double* scalars, double* x, double* s,
                     double* y, double* z) {
  const auto& r = H(h)->problem->last_trace().rows[row];
  double sc[16] = {double(r.iteration), double(r.type), r.error, r.cost,
                   r.infeasibility, r.complementarity, r.mu, r.delta, r.gamma,
                   r.alpha, r.alpha_max, r.alpha_z, double(r.factorizations),
                   double(r.solves), double(r.trials), r.t_end};
  std::memcpy(scalars, sc, sizeof(sc));
  auto cp = [](double* dst, const std::vector<double>& v) {
    if (dst && !v.empty()) std::memcpy(dst, v.data(), v.size() * 8);
  };
  cp(x, r.x);
  cp(s, r.s);
  cp(y, r.y);
  cp(z, r.z);
}

void slpbh_solution(void* h, double* x, double* s, double* y, double* z) {
  Handle* hd = H(h);
  auto& vars = hd->problem->decision_variables();
  if (x) {
    for (size_t i = 0; i < vars.size(); ++i) x[i] = vars[i].value();
  }
  auto cp = [](double* dst, const std::vector<double>& v) {
    if (dst && !v.empty()) std::memcpy(dst, v.data(), v.size() * 8);
  };
  cp(s, hd->problem->last_s());
  cp(y, hd->problem->last_y());
  cp(z, hd->problem->last_z());
}

/// 0 interior-point, 1 SQP, 2 Newton: the branch the last solve() took.
int slpbh_solver_kind(void* h) {
  return static_cast<int>(H(h)->problem->last_solver_kind());
}

double slpbh_loop_seconds(void* h) {
  return H(h)->problem->last_trace().loop_seconds;
}

/// out[8]: dim, nnz_kkt, nnz_l, nnz_l_stored, n_supernodes, n_levels,
/// max_front, etree_height
void slpbh_symbolic_stats(void* h, int64_t* out) {
  const auto& s = H(h)->problem->last_symbolic_stats();
  out[0] = s.dim;
  out[1] = s.nnz_kkt;
  out[2] = s.nnz_l;
  out[3] = s.nnz_l_stored;
  out[4] = s.n_supernodes;
  out[5] = s.n_levels;
  out[6] = s.max_front;
  out[7] = s.etree_height;
}

/// out[12]: the fields of slpb_counters in declaration order.
void slpbh_counters(void* h, int64_t* out) {
  const auto& c = H(h)->problem->last_counters();
  const int64_t v[12] = {c.kernel_launches, c.factorizations, c.solves,
                         c.evals_full, c.evals_values, c.tape_nodes,
                         c.program_bytes, c.n_clusters, c.n_program_classes,
                         c.h2d_bytes, c.d2h_bytes, c.factorizations_completed};
  std::memcpy(out, v, sizeof(v));
}

/// out[15]: total_ms[5], count[5] (sampled runs), launches[5] of slpb_timers.
void slpbh_timers(void* h, double* out) {
  const auto& t = H(h)->problem->last_timers();
  for (int i = 0; i < 5; ++i) {
    out[i] = t.total_ms[i];
    out[5 + i] = static_cast<double>(t.count[i]);
    out[10 + i] = static_cast<double>(t.launches[i]);
  }
}

/// out[9]: count[3], bytes[3], total_ms[3] of slpb_comm_stats (sharded solves).
void slpbh_comm_stats(void* h, double* out) {
  const auto& c = H(h)->problem->last_comm_stats();
  for (int i = 0; i < 3; ++i) {
    out[i] = static_cast<double>(c.count[i]);
    out[3 + i] = static_cast<double>(c.bytes[i]);
    out[6 + i] = c.total_ms[i];
  }
}

/// Builds the autodiff graphs, opens a device handle, uploads and finalises.
/// Returns the raw slpb_solver* (owned by this handle) so a test can drive the
/// C ABI of include/slpb.h directly, or NULL on failure.
void* slpbh_device_open(void* h, int device) {
  Handle* hd = H(h);
  try {
    hd->graphs = hd->problem->build_graphs();
    slp::detail::FlatProblem fp = hd->graphs->flatten();
    hd->device = std::make_unique<slp::DeviceHandle>(device);
    slp::Problem<double>::upload(hd->device->s, fp);
    return hd->device->s;
  } catch (const std::exception& e) {
    hd->error = e.what();
    hd->device.reset();
    return nullptr;
  }
}

void slpbh_device_close(void* h) { H(h)->device.reset(); }

}  // extern "C"
