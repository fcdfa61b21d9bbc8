// TEST INFRASTRUCTURE — host emulation of the device library's *logic*.
//
// Lets the CPU-only test tier (pytest -m "not gpu") exercise, without a GPU:
//   * the host DSL → tape flattening (sleipnir_b200/include/sleipnir/...),
//   * the tape → cluster-program compiler (csrc/compile.cpp),
//   * the kernels' per-cluster / per-front bodies (csrc/ad_core.hpp,
//     csrc/ldlt_core.hpp), run with ONE lane instead of a warp,
//   * the symbolic analysis (csrc/symbolic.cpp).
// It is built only by tests/emu/Makefile into tests/emu/libslpb_emu.so and is
// never linked into, loaded by or reachable from the product
// (sleipnir_b200/): the product library has no CPU path and returns
// SLPB_ERR_NO_DEVICE without a GPU.
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "ad_core.hpp"
#include "internal.hpp"
#include "kkt_core.hpp"
#include "ldlt_core.hpp"
#include "problems/problems.hpp"

namespace {

struct Emu {
  std::unique_ptr<slp::Problem<double>> problem;
  std::unique_ptr<slp::Problem<double>::Graphs> graphs;
  slp::detail::FlatProblem flat;
  slpb::Tape tape;
  slpb::RowSet rows[SLPB_OUT_COUNT];
  slpb::CompiledAD ad;
  std::string error;
  int n = 0, me = 0, mi = 0;
  bool fused = false;  // arithmetic mode of the factorisation (ldlt_core.hpp)
  // last evaluation
  std::vector<double> values, derivs;
  // linear algebra
  slpb::KktRecipe recipe;
  slpb::Symbolic sym;
  std::vector<double> Kval, panels, updates, D, uvecs, xperm;
  int shard_world = 1;  // > 1: emulate the multi-GPU sharded sweep
};

slpb::SymbolicView view_of(const slpb::Symbolic& S) {
  slpb::SymbolicView v{};
  v.dim = S.dim;
  v.n_super = S.n_super;
  v.super_first = S.super_first.data();
  v.front_dim = S.front_dim.data();
  v.rows_ptr = S.rows_ptr.data();
  v.rows_idx = S.rows_idx.data();
  v.panel_ptr = S.panel_ptr.data();
  v.update_ptr = S.update_ptr.data();
  v.child_ptr = S.child_ptr.data();
  v.child_idx = S.child_idx.data();
  v.rel_ptr = S.rel_ptr.data();
  v.rel_idx = S.rel_idx.data();
  v.asm_ptr = S.asm_ptr.data();
  v.asm_src = S.asm_src.data();
  v.asm_dst = S.asm_dst.data();
  v.col_is_primal = S.col_is_primal.data();
  v.perm = S.perm.data();
  return v;
}

template <int LC>
void run_task(const slpb::ProgramSet& ps, int t, const double* leaf,
              double* stage, double* scratch) {
  const uint32_t* P = ps.blob.data() + ps.prog_offset[ps.task_prog[t]];
  const uint32_t* B = ps.task_bindings.data() + ps.task_bind[t];
  // nthreads = LC: every lane owns all items of its own cluster, so the lanes
  // are independent and may run one after the other
  slpb::DirectStream stream{P};
  for (int tid = 0; tid < LC; ++tid) {
    slpb::ad_run_group<LC>(tid, LC, ps.task_count[t], P, stream, B, leaf,
                           stage, scratch, slpb::NoSync{});
  }
}

/// Runs the task plan the device would run (same tasks, same transposed
/// bindings, same lane count per task).
void run_programs(const slpb::ProgramSet& ps, const double* leaf,
                  double* stage) {
  std::vector<double> scratch(size_t(ps.max_smem / 8 + 1) * 32);
  for (size_t t = 0; t < ps.task_prog.size(); ++t) {
    const int ti = static_cast<int>(t);
    switch (ps.task_lanes[t]) {
      case 1: run_task<1>(ps, ti, leaf, stage, scratch.data()); break;
      case 2: run_task<2>(ps, ti, leaf, stage, scratch.data()); break;
      case 4: run_task<4>(ps, ti, leaf, stage, scratch.data()); break;
      case 8: run_task<8>(ps, ti, leaf, stage, scratch.data()); break;
      case 16: run_task<16>(ps, ti, leaf, stage, scratch.data()); break;
      default: run_task<32>(ps, ti, leaf, stage, scratch.data()); break;
    }
  }
}

void run_one_task(const slpb::ProgramSet& ps, int ti, const double* leaf,
                  double* stage, double* scratch) {
  switch (ps.task_lanes[ti]) {
    case 1: run_task<1>(ps, ti, leaf, stage, scratch); break;
    case 2: run_task<2>(ps, ti, leaf, stage, scratch); break;
    case 4: run_task<4>(ps, ti, leaf, stage, scratch); break;
    case 8: run_task<8>(ps, ti, leaf, stage, scratch); break;
    case 16: run_task<16>(ps, ti, leaf, stage, scratch); break;
    default: run_task<32>(ps, ti, leaf, stage, scratch); break;
  }
}

/// What `world` ranks would do with a ShardPlan: every rank sweeps only its
/// task ranges into its OWN copy of the stage, packs the slots it produced,
/// the packed buffers are concatenated (the all-gather) and unpacked into
/// rank 0's stage, which is returned in `stage`.
void run_programs_sharded(const slpb::ProgramSet& ps, int world,
                          const double* leaf, std::vector<double>& stage) {
  slpb::ShardPlan plan;
  slpb::build_shard_plan(ps, world, plan);
  std::vector<double> scratch(size_t(ps.max_smem / 8 + 1) * 32);
  std::vector<double> gathered(size_t(world) * plan.max_len, 0.0);
  const std::vector<double> init = stage;
  const size_t n_launch = ps.launches.size();
  for (int r = 0; r < world; ++r) {
    std::vector<double> mine = init;
    for (size_t L = 0; L < n_launch; ++L) {
      const int first = plan.first_task[L * world + r];
      const int cnt = plan.n_tasks[L * world + r];
      for (int t = first; t < first + cnt; ++t) {
        run_one_task(ps, t, leaf, mine.data(), scratch.data());
      }
    }
    for (int i = 0; i < plan.len[r]; ++i) {
      gathered[size_t(r) * plan.max_len + i] =
          mine[plan.slots[size_t(r) * plan.max_len + i]];
    }
    if (r == 0) stage = mine;
  }
  for (int r = 1; r < world; ++r) {
    for (int i = 0; i < plan.len[r]; ++i) {
      stage[plan.slots[size_t(r) * plan.max_len + i]] =
          gathered[size_t(r) * plan.max_len + i];
    }
  }
}

void run_gather(const slpb::Gather& g, const double* stage, double d_f,
                const double* d_c, std::vector<double>& out) {
  out.resize(g.n_entries());
  for (int e = 0; e < g.n_entries(); ++e) {
    out[e] = slpb::gather_entry(e, g.ptr.data(), g.src_idx.data(),
                                g.src_scale.data(), stage, d_f, d_c);
  }
}

}  // namespace

extern "C" {

void* emu_create(const char* name, int N, double p0, double p1) {
  auto e = std::make_unique<Emu>();
  try {
    e->problem = slpb_problems::make_problem(name, N, p0, p1);
  } catch (...) {
    return nullptr;
  }
  e->n = static_cast<int>(e->problem->decision_variables().size());
  e->me = static_cast<int>(e->problem->equality_constraints().size());
  e->mi = static_cast<int>(e->problem->inequality_constraints().size());
  e->graphs = e->problem->build_graphs();
  e->flat = e->graphs->flatten();
  const auto& f = e->flat;
  if (!slpb::ingest_tape(e->tape, f.n_nodes(), f.op.data(), f.lhs.data(),
                         f.rhs.data(), f.val.data(), (int)f.leaf_x.size(),
                         f.leaf_x.data(), (int)f.leaf_y.size(),
                         f.leaf_y.data(), (int)f.leaf_z.size(),
                         f.leaf_z.data(), e->error)) {
    return e.release();
  }
  for (int w = 0; w < SLPB_OUT_COUNT; ++w) {
    slpb_rowset view = f.rows[w].view();
    if (!slpb::ingest_rows(e->rows[w], e->tape, w, &view,
                           f.rows[w].const_val.data(), e->error)) {
      return e.release();
    }
  }
  if (!slpb::compile_autodiff(e->tape, e->rows, false, e->ad)) {
    e->error = e->ad.error;
    return e.release();
  }
  // the device's plan, with a deliberately small budget when asked (so that
  // tests also cover tasks of fewer than 32 lanes)
  int32_t budget = 200 * 1024;
  if (const char* env = std::getenv("SLPB_EMU_SMEM_BUDGET")) budget = std::atoi(env);
  if (!slpb::build_task_plan(e->ad.values, budget, e->error) ||
      !slpb::build_task_plan(e->ad.derivs, budget, e->error)) {
    return e.release();
  }
  return e.release();
}

void emu_destroy(void* h) { delete static_cast<Emu*>(h); }
/// Emulates the derivative sweep as `world` ranks would run it (ShardPlan).
void emu_set_shard_world(void* h, int world) {
  static_cast<Emu*>(h)->shard_world = world;
}
const char* emu_error(void* h) { return static_cast<Emu*>(h)->error.c_str(); }

void emu_dims(void* h, int* n, int* me, int* mi) {
  auto* e = static_cast<Emu*>(h);
  *n = e->n;
  *me = e->me;
  *mi = e->mi;
}

void emu_initial_guess(void* h, double* x) {
  auto* e = static_cast<Emu*>(h);
  auto& vars = e->problem->decision_variables();
  for (size_t i = 0; i < vars.size(); ++i) x[i] = vars[i].value();
}

/// out[0..11]: tape nodes, value clusters, value programs, value blob words,
/// deriv clusters, deriv programs, deriv blob words, max smem bytes, instr,
/// visits, contribs, deriv stage size
void emu_stats(void* h, int64_t* out) {
  auto* e = static_cast<Emu*>(h);
  const auto& a = e->ad;
  out[0] = e->tape.n_nodes;
  out[1] = (int64_t)a.values.cluster_prog.size();
  out[2] = (int64_t)a.values.prog_offset.size();
  out[3] = (int64_t)a.values.blob.size();
  out[4] = (int64_t)a.derivs.cluster_prog.size();
  out[5] = (int64_t)a.derivs.prog_offset.size();
  out[6] = (int64_t)a.derivs.blob.size();
  out[7] = std::max(a.values.max_smem, a.derivs.max_smem);
  out[8] = a.derivs.n_instr;
  out[9] = a.derivs.n_visits;
  out[10] = a.derivs.n_contribs;
  out[11] = a.deriv_stage_size;
}

/// Header words of program `prog` of the value (set 0) or derivative (set 1)
/// program set; returns the number of programs in the set.
int emu_program_header(void* h, int set, int prog, uint32_t* out24) {
  auto* e = static_cast<Emu*>(h);
  const slpb::ProgramSet& ps = set == 0 ? e->ad.values : e->ad.derivs;
  const int n = static_cast<int>(ps.prog_offset.size());
  if (prog >= 0 && prog < n) {
    std::memcpy(out24, ps.blob.data() + ps.prog_offset[prog], 24 * 4);
  }
  return n;
}

/// which: SLPB_OUT_A_E, SLPB_OUT_A_I, SLPB_OUT_H_C (pattern of H).
void emu_pattern(void* h, int which, int* rows, int* cols, int64_t* nnz,
                 int* colptr, int* rowidx) {
  auto* e = static_cast<Emu*>(h);
  const slpb::Pattern& p = which == SLPB_OUT_A_E   ? e->ad.A_e
                           : which == SLPB_OUT_A_I ? e->ad.A_i
                                                   : e->ad.H;
  *rows = p.rows;
  *cols = p.cols;
  *nnz = p.nnz();
  if (colptr) {
    std::memcpy(colptr, p.colptr.data(), p.colptr.size() * sizeof(int));
    if (p.nnz()) std::memcpy(rowidx, p.rowidx.data(), p.nnz() * sizeof(int));
  }
}

/// Evaluates everything at (x, y, z) with scaling (d_f, d_ce, d_ci). Outputs:
/// f (1), c_e (me), c_i (mi), g (n), A_e.val, A_i.val, H.val.
void emu_eval(void* h, const double* x, const double* y, const double* z,
              double d_f, const double* d_ce, const double* d_ci, double* f,
              double* c_e, double* c_i, double* g, double* ae, double* ai,
              double* hv) {
  auto* e = static_cast<Emu*>(h);
  const int n = e->n, me = e->me, mi = e->mi;
  std::vector<double> leaf(n + me + mi), d_c(me + mi);
  for (int i = 0; i < n; ++i) leaf[i] = x[i];
  for (int i = 0; i < me; ++i) {
    d_c[i] = d_ce[i];
    leaf[n + i] = d_ce[i] * y[i];
  }
  for (int i = 0; i < mi; ++i) {
    d_c[me + i] = d_ci[i];
    leaf[n + me + i] = d_ci[i] * z[i];
  }
  const auto& a = e->ad;
  std::vector<double> vstage = a.value_stage_init;
  run_programs(a.values, leaf.data(), vstage.data());
  run_gather(a.value_gather, vstage.data(), d_f, d_c.data(), e->values);
  std::vector<double> dstage = a.deriv_stage_init;
  if (e->shard_world > 1) {
    run_programs_sharded(a.derivs, e->shard_world, leaf.data(), dstage);
  } else {
    run_programs(a.derivs, leaf.data(), dstage.data());
  }
  run_gather(a.deriv_gather, dstage.data(), d_f, d_c.data(), e->derivs);
  *f = e->values[0];
  std::memcpy(c_e, e->values.data() + 1, me * 8);
  std::memcpy(c_i, e->values.data() + 1 + me, mi * 8);
  std::memcpy(g, e->derivs.data() + a.off_g, n * 8);
  if (a.A_e.nnz()) std::memcpy(ae, e->derivs.data() + a.off_ae, a.A_e.nnz() * 8);
  if (a.A_i.nnz()) std::memcpy(ai, e->derivs.data() + a.off_ai, a.A_i.nnz() * 8);
  if (a.H.nnz()) std::memcpy(hv, e->derivs.data() + a.off_h, a.H.nnz() * 8);
}


// ---- KKT assembly + multifrontal LDLT --------------------------------------

/// Builds the KKT recipe; returns nnz of the lower triangle.
int64_t emu_kkt_build(void* h) {
  auto* e = static_cast<Emu*>(h);
  slpb::build_kkt_recipe(e->n, e->me, e->ad.H, e->ad.A_e, e->ad.A_i, e->recipe);
  return e->recipe.K.nnz();
}

void emu_kkt_pattern(void* h, int* colptr, int* rowidx) {
  auto* e = static_cast<Emu*>(h);
  const auto& K = e->recipe.K;
  std::memcpy(colptr, K.colptr.data(), K.colptr.size() * sizeof(int));
  std::memcpy(rowidx, K.rowidx.data(), K.nnz() * sizeof(int));
}

/// Assembles lhs values (no δ/γ) from the last emu_eval and Σ = z/s.
void emu_kkt_assemble(void* h, const double* sigma, double* kval_out) {
  auto* e = static_cast<Emu*>(h);
  const auto& R = e->recipe;
  const auto& a = e->ad;
  e->Kval.resize(R.K.nnz());
  for (int64_t k = 0; k < R.K.nnz(); ++k) {
    e->Kval[k] = slpb::kkt_entry(
        static_cast<int>(k), R.h_idx.data(), R.ae_idx.data(), R.prod_ptr.data(),
        R.prod_a.data(), R.prod_b.data(), R.prod_row.data(),
        e->derivs.data() + a.off_h, e->derivs.data() + a.off_ae,
        e->derivs.data() + a.off_ai, sigma);
  }
  if (kval_out) std::memcpy(kval_out, e->Kval.data(), e->Kval.size() * 8);
}

/// Replaces the assembled lhs values (e.g. by the device's, so that the two
/// factorisations start from the same bits).
void emu_set_kkt_values(void* h, const double* kval) {
  auto* e = static_cast<Emu*>(h);
  e->Kval.assign(kval, kval + e->recipe.K.nnz());
}

/// out: dim, nnz_l, n_super, n_levels, max_front, etree_height, panel doubles,
/// update doubles. Returns 0 on success.
int emu_analyze(void* h, int ordering, const int* perm, int64_t* out) {
  auto* e = static_cast<Emu*>(h);
  if (!slpb::analyze_kkt(e->recipe.K, e->n, ordering, perm, e->sym, e->error)) {
    return -1;
  }
  const auto& S = e->sym;
  out[0] = S.dim;
  out[1] = S.nnz_l;
  out[2] = S.n_super;
  out[3] = S.n_levels;
  out[4] = S.max_front;
  out[5] = S.etree_height;
  out[6] = S.panel_size;
  out[7] = S.update_size;
  e->panels.assign(S.panel_size, 0.0);
  e->updates.assign(S.update_size, 0.0);
  e->D.assign(S.dim, 0.0);
  e->uvecs.assign(S.rel_ptr.back(), 0.0);
  e->xperm.assign(S.dim, 0.0);
  return 0;
}

/// Per front of the last analysis: order, own columns, level, parent.
int emu_fronts(void* h, int* F, int* np, int* level, int* parent) {
  auto* e = static_cast<Emu*>(h);
  const auto& S = e->sym;
  for (int s = 0; s < S.n_super; ++s) {
    F[s] = S.front_dim[s];
    np[s] = S.super_first[s + 1] - S.super_first[s];
    level[s] = S.super_level[s];
    parent[s] = S.super_parent[s];
  }
  return S.n_super;
}

/// Multi-GPU partition of the assembly tree of the last analysis
/// (build_tree_shard): owner per front (−1: replicated top), and per rank the
/// work it owns; work[world] = work of the top. Returns the number of top fronts.
int emu_tree_shard(void* h, int world, int* owner, double* work) {
  auto* e = static_cast<Emu*>(h);
  slpb::TreeShard sh;
  slpb::build_tree_shard(e->sym, world, sh);
  for (int s = 0; s < e->sym.n_super; ++s) owner[s] = sh.owner[s];
  for (int r = 0; r < world; ++r) work[r] = sh.rank_work[r];
  work[world] = sh.top_work;
  // self-checks of the lists the device side relies on
  size_t listed = sh.top_order.size();
  for (const auto& o : sh.rank_order) listed += o.size();
  if (listed != static_cast<size_t>(e->sym.n_super)) return -1;
  for (int s = 0; s < e->sym.n_super; ++s) {
    const int p = e->sym.super_parent[s];
    // the top is closed upwards; a subtree has one owner
    if (sh.owner[s] < 0 && p >= 0 && sh.owner[p] >= 0) return -2;
    if (sh.owner[s] >= 0 && p >= 0 && sh.owner[p] >= 0 && sh.owner[p] != sh.owner[s]) return -3;
  }
  return static_cast<int>(sh.top_order.size());
}

/// Raw output of the product's approximate-minimum-degree ordering
/// (csrc/amd.cpp) on a lower-triangular pattern, before analyze_kkt folds the
/// elimination-tree postorder into it.
void emu_order_amd(int n, const int* colptr, const int* rowidx, int* perm) {
  slpb::Pattern L;
  L.rows = L.cols = n;
  L.colptr.assign(colptr, colptr + n + 1);
  L.rowidx.assign(rowidx, rowidx + colptr[n]);
  const std::vector<int32_t> p = slpb::order_amd(L);
  std::memcpy(perm, p.data(), n * sizeof(int));
}

void emu_get_perm(void* h, int* perm) {
  auto* e = static_cast<Emu*>(h);
  std::memcpy(perm, e->sym.perm.data(), e->sym.dim * sizeof(int));
}

/// info: n_pos, n_neg, n_zero, zero_pivot; returns min |D|.
void emu_set_fused(void* h, int fused) { static_cast<Emu*>(h)->fused = fused != 0; }

double emu_factor(void* h, double delta, double gamma, int* info, double* D) {
  auto* e = static_cast<Emu*>(h);
  const auto& S = e->sym;
  slpb::SymbolicView V = view_of(S);
  std::vector<double> W(size_t(S.max_front) * S.max_front), lcol(S.max_front);
  int tot[4] = {0, 0, 0, 0};
  double min_abs = INFINITY;
  for (int L = 0; L < S.n_levels; ++L) {
    for (int k = S.level_ptr[L]; k < S.level_ptr[L + 1]; ++k) {
      int ls[6];
      slpb::ldlt_factor_front<1>(0, S.level_supers[k], V, e->Kval.data(), delta,
                                 gamma, e->panels.data(), e->updates.data(),
                                 e->D.data(), W.data(), lcol.data(), ls,
                                 slpb::NoSync{}, e->fused);
      for (int i = 0; i < 4; ++i) tot[i] += ls[i];
      unsigned long long bits = (unsigned long long)(unsigned)ls[4] |
                                ((unsigned long long)(unsigned)ls[5] << 32);
      double a;
      std::memcpy(&a, &bits, 8);
      min_abs = std::fmin(min_abs, a);
    }
  }
  std::memcpy(info, tot, sizeof(tot));
  if (D) std::memcpy(D, e->D.data(), S.dim * 8);
  return min_abs;
}

void emu_solve(void* h, const double* rhs, double* x) {
  auto* e = static_cast<Emu*>(h);
  const auto& S = e->sym;
  slpb::SymbolicView V = view_of(S);
  std::vector<double> w(S.max_front);
  for (int L = 0; L < S.n_levels; ++L) {
    for (int k = S.level_ptr[L]; k < S.level_ptr[L + 1]; ++k) {
      slpb::ldlt_forward_front<1>(0, S.level_supers[k], V, e->panels.data(),
                                  rhs, e->xperm.data(), e->uvecs.data(),
                                  w.data(), slpb::NoSync{});
    }
  }
  for (int L = S.n_levels - 1; L >= 0; --L) {
    for (int k = S.level_ptr[L]; k < S.level_ptr[L + 1]; ++k) {
      slpb::ldlt_backward_front<1>(0, S.level_supers[k], V, e->panels.data(),
                                   e->D.data(), e->xperm.data(), w.data(),
                                   slpb::NoSync{});
    }
  }
  for (int k = 0; k < S.dim; ++k) x[S.perm[k]] = e->xperm[k];
}

}  // extern "C"
