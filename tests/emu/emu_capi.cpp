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
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "ad_core.hpp"
#include "internal.hpp"
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
  // last evaluation
  std::vector<double> values, derivs;
};

void run_programs(const slpb::ProgramSet& ps, const double* leaf,
                  double* stage) {
  std::vector<double> scratch(ps.max_smem / 8 + 1);
  for (size_t c = 0; c < ps.cluster_prog.size(); ++c) {
    const uint32_t* P = ps.blob.data() + ps.prog_offset[ps.cluster_prog[c]];
    const uint32_t* B = ps.bindings.data() + ps.cluster_bind[c];
    slpb::ad_run_cluster<1>(0, P, B, leaf, stage, scratch.data(),
                            slpb::NoSync{});
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
  }
  return e.release();
}

void emu_destroy(void* h) { delete static_cast<Emu*>(h); }
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
  run_programs(a.derivs, leaf.data(), dstage.data());
  run_gather(a.deriv_gather, dstage.data(), d_f, d_c.data(), e->derivs);
  *f = e->values[0];
  std::memcpy(c_e, e->values.data() + 1, me * 8);
  std::memcpy(c_i, e->values.data() + 1 + me, mi * 8);
  std::memcpy(g, e->derivs.data() + a.off_g, n * 8);
  if (a.A_e.nnz()) std::memcpy(ae, e->derivs.data() + a.off_ae, a.A_e.nnz() * 8);
  if (a.A_i.nnz()) std::memcpy(ai, e->derivs.data() + a.off_ai, a.A_i.nnz() * 8);
  if (a.H.nnz()) std::memcpy(hv, e->derivs.data() + a.off_h, a.H.nnz() * 8);
}

}  // extern "C"
