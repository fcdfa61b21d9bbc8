// slpb_group — dynamic batching of the linear algebra of concurrent solves that
// share ONE KKT pattern (included at the end of slpb.cu, after batch.cuh).
//
// Reference surface: slp::multistart (optimization/multistart.hpp:44-73) runs
// the same problem from many initial guesses, every start on its own thread.
// On the device every start keeps its own handle, stream and iterate; what they
// share is the structure of lhs (interior_point.hpp:426-440). Members of a group
// hand their factorisations and triangular solves to ONE batched launch
// (batch.cuh: lane = instance): a member that reaches slpb_factor / slpb_solve
// parks its request and waits; when every active member has parked one (or
// has left), the last arriver runs the round — scatter of the members' lhs
// values / right-hand sides into the SoA batch, one k_batch_factor and/or one
// k_batch_solve for all of them, results handed back — and wakes the others.
// The per-instance arithmetic is that of the single-instance kernels, so a
// grouped solve reproduces the solitary one bit for bit, whatever the others do.
#pragma once

#include <condition_variable>

struct slpb_group {
  int device = 0;
  int expected = 0;       // members that will join (or be abandoned)
  std::mutex mu;
  std::condition_variable cv;
  slpb_batch* batch = nullptr;  // created from the first member's analysis
  // signature of the first member: later members must match it
  int64_t nK = 0;
  int32_t dim = 0;
  std::vector<int32_t> perm, k_colptr, k_rowidx;

  struct Request {
    int kind = 0;  // 0 none, 1 factor, 2 solve
    int nv = 1;
    double delta[2] = {0, 0}, gamma[2] = {0, 0};
    slpb_factor_info info[2] = {};
    int sel = 0;
    int rc = SLPB_OK;
  };
  struct Member {
    slpb_solver* S = nullptr;
    bool active = false;   // takes part in the rounds
    Request req;
  };
  std::vector<Member> members;   // slot m owns batch instances 2m, 2m + 1
  int joined = 0, abandoned = 0;
  int active = 0, parked = 0;
  uint64_t round = 0;
  int64_t rounds_run = 0, requests_served = 0;
  std::vector<double> cur_delta, cur_gamma;  // per batch instance
  // where the wall time of the rounds goes (host clock, seconds)
  double t_scatter = 0, t_factor = 0, t_solve = 0, t_gather = 0, t_quorum = 0;
  int64_t n_factor_rounds = 0, n_solve_rounds = 0, n_factor_req = 0, n_solve_req = 0;
  std::chrono::steady_clock::time_point last_round_end = std::chrono::steady_clock::now();
  std::string error;
};

namespace slpb {

/// Runs one round for every parked request (called with the lock held by the
/// thread that completed the quorum). Everything is enqueued on the batch's
/// stream; members synchronised their own streams before parking.
inline void group_run_round(slpb_group* G) {
  slpb_batch* B = G->batch;
  const size_t lanes = size_t(B->groups) * 32;
  bool any_factor = false, any_solve = false;
  for (auto& m : G->members) {
    any_factor |= m.req.kind == 1;
    any_solve |= m.req.kind == 2;
  }
  auto fail_all = [&](int rc) {
    for (auto& m : G->members) {
      if (m.req.kind != 0) {
        m.req.rc = rc;
        m.req.kind = 0;
      }
    }
  };
  cudaSetDevice(G->device);
  using clk = std::chrono::steady_clock;
  auto secs = [](clk::time_point a, clk::time_point b) {
    return std::chrono::duration<double>(b - a).count();
  };
  const auto t_begin = clk::now();
  G->t_quorum += secs(G->last_round_end, t_begin);
  struct Stamp {
    slpb_group* G;
    ~Stamp() { G->last_round_end = clk::now(); }
  } stamp{G};
  if (any_factor) {
    ++G->n_factor_rounds;
    const auto t0 = clk::now();
    // A launch factors EVERY slot: the slots without a request keep the
    // regularisation of their last one, so that their factor — which their
    // owner may still be about to solve with — is reproduced, not clobbered.
    if (G->cur_delta.size() != size_t(B->batch)) {
      G->cur_delta.assign(B->batch, 1.0);
      G->cur_gamma.assign(B->batch, 1.0);
    }
    std::vector<double>& d = G->cur_delta;
    std::vector<double>& g = G->cur_gamma;
    for (size_t i = 0; i < G->members.size(); ++i) {
      auto& m = G->members[i];
      if (m.req.kind != 1) continue;
      for (int v = 0; v < m.req.nv; ++v) {
        const int inst = static_cast<int>(2 * i + v);
        const int gidx = inst / 32, l = inst % 32;
        k_batch_scatter<<<blocks_for(B->nK, 256), 256, 0, B->stream>>>(
            m.S->Kval.p, B->nK, B->Kb.p + int64_t(gidx) * B->nK * 32 + l);
        d[inst] = m.req.delta[v];
        g[inst] = m.req.gamma[v];
      }
    }
    std::vector<slpb_factor_info> info(B->batch);
    const auto t1 = clk::now();
    G->t_scatter += secs(t0, t1);
    const int rc = slpb_batch_factor(B, d.data(), g.data(), info.data());
    G->t_factor += secs(t1, clk::now());
    if (rc != SLPB_OK) {
      G->error = B->error;
      fail_all(rc);
      return;
    }
    for (size_t i = 0; i < G->members.size(); ++i) {
      auto& m = G->members[i];
      if (m.req.kind != 1) continue;
      G->n_factor_req += m.req.nv;
      for (int v = 0; v < m.req.nv; ++v) m.req.info[v] = info[2 * i + v];
      m.req.rc = SLPB_OK;
      m.req.kind = 0;  // served: the answer stays in the slot for its owner
    }
  }
  if (any_solve) {
    ++G->n_solve_rounds;
    const auto t0 = clk::now();
    for (size_t i = 0; i < G->members.size(); ++i) {
      auto& m = G->members[i];
      if (m.req.kind != 2) continue;
      const int inst = static_cast<int>(2 * i + m.req.sel);
      const int gidx = inst / 32, l = inst % 32;
      k_batch_scatter<<<blocks_for(B->dim, 256), 256, 0, B->stream>>>(
          m.S->rhs.p, B->dim, B->rhs.p + int64_t(gidx) * B->dim * 32 + l);
    }
    const auto t1 = clk::now();
    G->t_scatter += secs(t0, t1);
    const int rc = slpb_batch_solve(B);
    const auto t2 = clk::now();
    G->t_solve += secs(t1, t2);
    if (rc != SLPB_OK) {
      G->error = B->error;
      fail_all(rc);
      return;
    }
    for (size_t i = 0; i < G->members.size(); ++i) {
      auto& m = G->members[i];
      if (m.req.kind != 2) continue;
      const int inst = static_cast<int>(2 * i + m.req.sel);
      const int gidx = inst / 32, l = inst % 32;
      k_batch_gather<<<blocks_for(B->dim, 256), 256, 0, B->stream>>>(
          B->sol.p + int64_t(gidx) * B->dim * 32 + l, B->dim, m.S->sol.p);
      m.req.rc = SLPB_OK;
    }
    const bool ok = cudaStreamSynchronize(B->stream) == cudaSuccess;
    G->t_gather += secs(t2, clk::now());
    for (auto& m : G->members) {
      if (m.req.kind != 2) continue;
      ++G->n_solve_req;
      if (!ok) m.req.rc = SLPB_ERR_CUDA;
      m.req.kind = 0;
    }
  }
  (void)lanes;
  ++G->rounds_run;
}

/// Parks `req` for member `slot` and returns once it has been served.
inline int group_submit(slpb_group* G, int slot, slpb_group::Request& req) {
  std::unique_lock<std::mutex> lock{G->mu};
  auto& m = G->members[slot];
  m.req = req;
  ++G->parked;
  ++G->requests_served;
  if (G->parked >= G->active) {
    group_run_round(G);
    req = m.req;
    // every parked member finds its answer in its own slot
    G->parked = 0;
    ++G->round;
    G->cv.notify_all();
    return req.rc;
  }
  const uint64_t r = G->round;
  G->cv.wait(lock, [&] { return G->round != r; });
  req = m.req;
  return req.rc;
}

/// A member stops taking part (finished, or busy elsewhere for long): whoever
/// is parked must not wait for it.
inline void group_deactivate(slpb_group* G, int slot) {
  std::unique_lock<std::mutex> lock{G->mu};
  auto& m = G->members[slot];
  if (!m.active) return;
  m.active = false;
  --G->active;
  if (G->parked > 0 && G->parked >= G->active) {
    group_run_round(G);
    G->parked = 0;
    ++G->round;
    G->cv.notify_all();
  }
}

int group_check_factor(slpb_solver* S, int nv, const double* delta,
                       const double* gamma, slpb_factor_info* info);

int group_factor(slpb_solver* S, int nv, const double* delta,
                        const double* gamma, slpb_factor_info* info) {
  slpb_group::Request req;
  req.kind = 1;
  req.nv = nv;
  for (int v = 0; v < nv; ++v) {
    req.delta[v] = delta[v];
    req.gamma[v] = gamma[v];
  }
  const int rc = group_submit(S->group, S->group_slot, req);
  if (rc != SLPB_OK) return fail(S, rc, "batched factorisation: " + S->group->error);
  for (int v = 0; v < nv; ++v) {
    info[v] = req.info[v];
    if (!info[v].zero_pivot) ++S->counters.factorizations_completed;
  }
  S->counters.factorizations += nv;
  S->fwd_valid[0] = S->fwd_valid[1] = false;
  S->factor_sel = 0;
  if (std::getenv("SLPB_GROUP_CHECK")) {
    // debugging aid: the member's own kernels on the same system must agree
    // with the batched launch bit for bit
    slpb_group* G = S->group;
    S->group = nullptr;
    slpb_factor_info own[2]{};
    const bool keep_rhs = S->rhs_ready;
    S->rhs_ready = false;
    const int rc2 = group_check_factor(S, nv, delta, gamma, own);
    S->rhs_ready = keep_rhs;
    S->group = G;
    if (rc2 != SLPB_OK) return rc2;
    std::vector<double> a(S->dim), b(S->dim);
    for (int v = 0; v < nv; ++v) {
      bool same = own[v].n_pos == info[v].n_pos && own[v].n_neg == info[v].n_neg &&
                  own[v].n_zero == info[v].n_zero &&
                  own[v].zero_pivot == info[v].zero_pivot;
      if (!own[v].zero_pivot) same = same && own[v].min_abs_d == info[v].min_abs_d;
      int diff = 0;
      if (!own[v].zero_pivot) {
        cudaMemcpy(a.data(), S->D.p + size_t(v) * S->dim, S->dim * 8,
                   cudaMemcpyDeviceToHost);
        {
          std::unique_lock<std::mutex> lock{G->mu};
          slpb_batch_get(G->batch, 2 * S->group_slot + v, SLPB_BATCH_D, b.data());
        }
        for (int i = 0; i < S->dim; ++i) diff += std::memcmp(&a[i], &b[i], 8) != 0;
      }
      if (!same || diff) {
        std::fprintf(stderr,
                     "[slpb group check] slot %d variant %d (delta %.3g gamma %.3g): "
                     "info own (%d %d %d %d %.17g) batch (%d %d %d %d %.17g), %d D entries differ\n",
                     S->group_slot, v, delta[v], gamma[v], own[v].n_pos, own[v].n_neg,
                     own[v].n_zero, own[v].zero_pivot, own[v].min_abs_d, info[v].n_pos,
                     info[v].n_neg, info[v].n_zero, info[v].zero_pivot,
                     info[v].min_abs_d, diff);
      }
    }
  }
  return SLPB_OK;
}

int group_solve(slpb_solver* S) {
  slpb_group::Request req;
  req.kind = 2;
  req.sel = S->factor_sel;
  const int rc = group_submit(S->group, S->group_slot, req);
  if (rc != SLPB_OK) return fail(S, rc, "batched solve: " + S->group->error);
  ++S->counters.solves;
  return SLPB_OK;
}

int group_check_factor(slpb_solver* S, int nv, const double* delta,
                       const double* gamma, slpb_factor_info* info) {
  return nv == 2 ? slpb_factor_pair(S, delta, gamma, 0, info)
                 : slpb_factor(S, delta[0], gamma[0], 0, info);
}

int group_download_d(slpb_solver* S, double* dst) {
  slpb_group* G = S->group;
  std::unique_lock<std::mutex> lock{G->mu};
  return slpb_batch_get(G->batch, 2 * S->group_slot + S->factor_sel,
                        SLPB_BATCH_D, dst);
}

}  // namespace slpb

extern "C" {

int slpb_group_create(int device, int32_t expected_members, slpb_group** out) {
  if (!out || expected_members < 1) return SLPB_ERR_ARGUMENT;
  auto G = std::make_unique<slpb_group>();
  G->device = device;
  G->expected = expected_members;
  G->members.resize(expected_members);
  *out = G.release();
  return SLPB_OK;
}

void slpb_group_destroy(slpb_group* G) {
  if (!G) return;
  if (std::getenv("SLPB_GROUP_TIMING")) {
    std::fprintf(stderr,
                 "[slpb group] %d members: %lld factor rounds (%lld requests), "
                 "%lld solve rounds (%lld requests); host seconds: scatter %.3f "
                 "factor %.3f solve %.3f gather %.3f, waiting for the quorum %.3f\n",
                 G->joined, (long long)G->n_factor_rounds, (long long)G->n_factor_req,
                 (long long)G->n_solve_rounds, (long long)G->n_solve_req,
                 G->t_scatter, G->t_factor, G->t_solve, G->t_gather, G->t_quorum);
  }
  if (G->batch) slpb_batch_destroy(G->batch);
  delete G;
}

int slpb_group_join(slpb_group* G, slpb_solver* S) {
  using namespace slpb;
  if (!G || !S) return SLPB_ERR_ARGUMENT;
  std::unique_lock<std::mutex> lock{G->mu};
  if (G->joined + G->abandoned >= G->expected) {
    return fail(S, SLPB_ERR_STATE, "slpb_group_join: the group is full");
  }
  // (every refusal below counts as "abandoned" for the start gate)
  if (!S->analyzed || !S->use_tree || S->world > 1 || S->device != G->device ||
      S->me + S->mi == 0) {
    ++G->abandoned;
    G->cv.notify_all();
    return fail(S, SLPB_ERR_STATE,
                "slpb_group_join: needs an analysed, constrained single-GPU "
                "solver with fronts of order <= 32 on the group's device");
  }
  if (G->batch == nullptr) {
    slpb_batch* b = nullptr;
    const int rc = slpb_batch_create(S, 2 * G->expected, &b);
    if (rc != SLPB_OK) {
      ++G->abandoned;
      G->cv.notify_all();
      return rc;
    }
    b->counters = nullptr;  // the first member may leave before the group
    G->batch = b;
    G->nK = S->recipe.K.nnz();
    G->dim = S->sym.dim;
    G->perm = S->sym.perm;
    G->k_colptr = S->recipe.K.colptr;
    G->k_rowidx = S->recipe.K.rowidx;
  } else if (S->recipe.K.nnz() != G->nK || S->sym.dim != G->dim ||
             S->sym.perm != G->perm || S->recipe.K.colptr != G->k_colptr ||
             S->recipe.K.rowidx != G->k_rowidx) {
    ++G->abandoned;  // counts towards the quorum of the start gate
    G->cv.notify_all();
    return fail(S, SLPB_ERR_ARGUMENT,
                "slpb_group_join: this solver's KKT pattern or elimination "
                "order differs from the group's");
  }
  const int slot = G->joined++;
  G->members[slot].S = S;
  G->members[slot].active = true;
  ++G->active;
  S->group = G;
  S->group_slot = slot;
  // start gate: the rounds begin when everybody is here (a member that starts
  // early would run its first rounds alone, batch of one)
  G->cv.notify_all();
  G->cv.wait(lock, [&] { return G->joined + G->abandoned >= G->expected; });
  return SLPB_OK;
}

int slpb_group_abandon(slpb_group* G) {
  if (!G) return SLPB_ERR_ARGUMENT;
  std::unique_lock<std::mutex> lock{G->mu};
  ++G->abandoned;
  G->cv.notify_all();
  return SLPB_OK;
}

int slpb_group_pause(slpb_solver* S) {
  if (!S) return SLPB_ERR_ARGUMENT;
  if (S->group) slpb::group_deactivate(S->group, S->group_slot);
  return SLPB_OK;
}

int slpb_group_resume(slpb_solver* S) {
  if (!S) return SLPB_ERR_ARGUMENT;
  if (!S->group) return SLPB_OK;
  slpb_group* G = S->group;
  std::unique_lock<std::mutex> lock{G->mu};
  auto& m = G->members[S->group_slot];
  if (!m.active) {
    m.active = true;
    ++G->active;
  }
  return SLPB_OK;
}

int slpb_group_leave(slpb_solver* S) {
  if (!S) return SLPB_ERR_ARGUMENT;
  if (!S->group) return SLPB_OK;
  slpb::group_deactivate(S->group, S->group_slot);
  {
    std::unique_lock<std::mutex> lock{S->group->mu};
    S->group->members[S->group_slot].S = nullptr;
  }
  S->group = nullptr;
  S->group_slot = -1;
  return SLPB_OK;
}

int slpb_group_stats(slpb_group* G, int64_t* rounds, int64_t* requests) {
  if (!G) return SLPB_ERR_ARGUMENT;
  std::unique_lock<std::mutex> lock{G->mu};
  if (rounds) *rounds = G->rounds_run;
  if (requests) *requests = G->requests_served;
  return SLPB_OK;
}

}  // extern "C"
