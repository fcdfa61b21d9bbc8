// Symbolic analysis of the reduced KKT system (host, once per solve).
//
// Replaces what Eigen::SimplicialLDLT::analyzePattern does for the reference
// (solver/util/sparse_regularized_ldlt.hpp:69-72: fill-reducing ordering +
// elimination tree + column counts), and additionally produces what a GPU
// factorisation needs: a supernodal assembly tree with dense fronts, a level
// schedule, and scatter maps from the static KKT pattern into the fronts.
//
// Ordering. The reference uses AMD, under which a direct-transcription KKT
// matrix has an elimination tree of height O(N) — a chain of ~2N dependent
// pivots that no GPU can hide. The default here is a nested-dissection order
// built from breadth-first level sets: a time-banded graph is split at a small
// vertex separator near the middle, recursively, which gives a tree of depth
// O(log N) whose leaves are independent stage blocks. Any permutation can be
// supplied instead (SLPB_ORDER_CUSTOM) — the parity tests hand the SAME
// permutation to the CPU oracle, because the regularisation decisions of the
// reference (min |D_ii| ≥ 1e-4, exact-zero pivots) are ordering-dependent.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <numeric>
#include <queue>

#include "internal.hpp"

namespace slpb {

void build_kkt_recipe(int32_t n, int32_t me, const Pattern& H,
                      const Pattern& A_e, const Pattern& A_i, KktRecipe& out) {
  out = KktRecipe{};
  const int32_t dim = n + me;
  // A_i by rows: (col, value index), columns ascending
  const int32_t mi = A_i.rows;
  std::vector<int32_t> rptr(mi + 1, 0);
  for (int32_t r : A_i.rowidx) ++rptr[r + 1];
  for (int32_t i = 0; i < mi; ++i) rptr[i + 1] += rptr[i];
  std::vector<int32_t> rcol(A_i.nnz()), ridx(A_i.nnz());
  {
    std::vector<int32_t> nxt(rptr.begin(), rptr.end() - 1);
    for (int32_t c = 0; c < A_i.cols; ++c) {
      for (int32_t k = A_i.colptr[c]; k < A_i.colptr[c + 1]; ++k) {
        const int32_t q = nxt[A_i.rowidx[k]]++;
        rcol[q] = c;
        ridx[q] = k;
      }
    }
  }
  // product terms keyed by (col, row) of K; generated with i ascending so that
  // a stable sort keeps Eigen's accumulation order (A_iᵀΣ)·A_i
  struct Term {
    int32_t c, r, a, b, i;
  };
  std::vector<Term> terms;
  for (int32_t i = 0; i < mi; ++i) {
    for (int32_t p = rptr[i]; p < rptr[i + 1]; ++p) {
      for (int32_t q = p; q < rptr[i + 1]; ++q) {
        // K(r, c) with r = rcol[q] ≥ c = rcol[p]
        terms.push_back({rcol[p], rcol[q], ridx[q], ridx[p], i});
      }
    }
  }
  std::stable_sort(terms.begin(), terms.end(), [](const Term& x, const Term& y) {
    return x.c != y.c ? x.c < y.c : x.r < y.r;
  });

  Pattern& K = out.K;
  K.rows = K.cols = dim;
  K.colptr.assign(dim + 1, 0);
  out.diag_idx.assign(dim, -1);
  size_t t = 0;
  std::vector<int32_t> rows;  // merge scratch
  for (int32_t c = 0; c < dim; ++c) {
    if (c < n) {
      int32_t hk = H.colptr[c];
      const int32_t hend = H.colptr[c + 1];
      size_t tk = t;
      while (tk < terms.size() && terms[tk].c == c) ++tk;
      bool diag_done = false;
      // three-way merge over rows < n: H, product terms, forced diagonal
      while (hk < hend || t < tk || !diag_done) {
        int32_t r = dim;
        if (hk < hend) r = std::min(r, H.rowidx[hk]);
        if (t < tk) r = std::min(r, terms[t].r);
        if (!diag_done) r = std::min(r, c);
        const int32_t e = static_cast<int32_t>(K.rowidx.size());
        K.rowidx.push_back(r);
        out.h_idx.push_back(-1);
        out.ae_idx.push_back(-1);
        out.prod_ptr.push_back(static_cast<int32_t>(out.prod_a.size()));
        if (hk < hend && H.rowidx[hk] == r) out.h_idx.back() = hk++;
        while (t < tk && terms[t].r == r) {
          out.prod_a.push_back(terms[t].a);
          out.prod_b.push_back(terms[t].b);
          out.prod_row.push_back(terms[t].i);
          ++t;
        }
        if (r == c) {
          diag_done = true;
          out.diag_idx[c] = e;
        }
      }
      for (int32_t k = A_e.colptr[c]; k < A_e.colptr[c + 1]; ++k) {
        K.rowidx.push_back(n + A_e.rowidx[k]);
        out.h_idx.push_back(-1);
        out.ae_idx.push_back(k);
        out.prod_ptr.push_back(static_cast<int32_t>(out.prod_a.size()));
      }
    } else {
      out.diag_idx[c] = static_cast<int32_t>(K.rowidx.size());
      K.rowidx.push_back(c);
      out.h_idx.push_back(-1);
      out.ae_idx.push_back(-1);
      out.prod_ptr.push_back(static_cast<int32_t>(out.prod_a.size()));
    }
    K.colptr[c + 1] = static_cast<int32_t>(K.rowidx.size());
  }
  out.prod_ptr.push_back(static_cast<int32_t>(out.prod_a.size()));
}

namespace {

/// Full symmetric adjacency (no diagonal) of a lower-triangular pattern.
void symmetric_adjacency(const Pattern& L, std::vector<int32_t>& ptr,
                         std::vector<int32_t>& idx) {
  const int32_t n = L.cols;
  ptr.assign(n + 1, 0);
  for (int32_t c = 0; c < n; ++c) {
    for (int32_t k = L.colptr[c]; k < L.colptr[c + 1]; ++k) {
      const int32_t r = L.rowidx[k];
      if (r == c) continue;
      ++ptr[c + 1];
      ++ptr[r + 1];
    }
  }
  for (int32_t i = 0; i < n; ++i) ptr[i + 1] += ptr[i];
  idx.assign(ptr[n], 0);
  std::vector<int32_t> nxt(ptr.begin(), ptr.end() - 1);
  for (int32_t c = 0; c < n; ++c) {
    for (int32_t k = L.colptr[c]; k < L.colptr[c + 1]; ++k) {
      const int32_t r = L.rowidx[k];
      if (r == c) continue;
      idx[nxt[c]++] = r;
      idx[nxt[r]++] = c;
    }
  }
}

struct NestedDissection {
  const std::vector<int32_t>& ptr;
  const std::vector<int32_t>& idx;
  int32_t leaf_size;
  std::vector<int32_t> part;   // current sub-graph id of each vertex
  std::vector<int32_t> dist;   // BFS level, −1 idle
  std::vector<int32_t> order;  // output: elimination order
  int32_t next_part = 1;

  NestedDissection(const std::vector<int32_t>& p, const std::vector<int32_t>& i,
                   int32_t leaf)
      : ptr{p}, idx{i}, leaf_size{leaf}, part(p.size() - 1, 0),
        dist(p.size() - 1, -1) {}

  /// BFS inside sub-graph `pid` from `start`; returns vertices level by level.
  void bfs(int32_t start, int32_t pid, std::vector<int32_t>& out,
           std::vector<int32_t>& level_ptr) {
    out.clear();
    level_ptr.assign(1, 0);
    out.push_back(start);
    dist[start] = 0;
    size_t head = 0;
    int32_t cur = 0;
    while (head < out.size()) {
      const int32_t u = out[head];
      if (dist[u] != cur) {
        level_ptr.push_back(static_cast<int32_t>(head));
        cur = dist[u];
      }
      ++head;
      for (int32_t k = ptr[u]; k < ptr[u + 1]; ++k) {
        const int32_t w = idx[k];
        if (part[w] == pid && dist[w] < 0) {
          dist[w] = dist[u] + 1;
          out.push_back(w);
        }
      }
    }
    level_ptr.push_back(static_cast<int32_t>(out.size()));
    for (int32_t v : out) dist[v] = -1;
  }

  void run() {
    const int32_t n = static_cast<int32_t>(part.size());
    // explicit work stack of (vertex list) to avoid deep recursion; separators
    // must be emitted AFTER both halves, so push a marker frame for them
    struct Frame {
      std::vector<int32_t> verts;
      bool emit_only;
    };
    std::vector<Frame> stack;
    {
      std::vector<int32_t> all(n);
      std::iota(all.begin(), all.end(), 0);
      stack.push_back({std::move(all), false});
    }
    std::vector<int32_t> comp, lvl, bfs_out;
    while (!stack.empty()) {
      Frame f = std::move(stack.back());
      stack.pop_back();
      if (f.emit_only) {
        order.insert(order.end(), f.verts.begin(), f.verts.end());
        continue;
      }
      // split into connected components
      const int32_t pid = next_part++;
      for (int32_t v : f.verts) part[v] = pid;
      std::vector<std::vector<int32_t>> comps;
      for (int32_t v : f.verts) {
        if (part[v] != pid) continue;
        bfs(v, pid, bfs_out, lvl);
        const int32_t cid = next_part++;
        for (int32_t w : bfs_out) part[w] = cid;
        comps.push_back(bfs_out);
        (void)cid;
      }
      // process components in reverse so the first one is emitted first
      for (auto it = comps.rbegin(); it != comps.rend(); ++it) {
        std::vector<int32_t>& cv = *it;
        if (static_cast<int32_t>(cv.size()) <= leaf_size) {
          stack.push_back({std::move(cv), true});
          continue;
        }
        const int32_t cid = part[cv[0]];
        // pseudo-peripheral start: repeat BFS from a min-degree vertex of the
        // last level
        int32_t start = cv[0];
        for (int pass = 0; pass < 3; ++pass) {
          bfs(start, cid, bfs_out, lvl);
          const int32_t lb = lvl[lvl.size() - 2], le = lvl.back();
          int32_t best = bfs_out[lb];
          for (int32_t k = lb; k < le; ++k) {
            const int32_t v = bfs_out[k];
            if (ptr[v + 1] - ptr[v] < ptr[best + 1] - ptr[best]) best = v;
          }
          start = best;
        }
        bfs(start, cid, bfs_out, lvl);
        const int32_t n_levels = static_cast<int32_t>(lvl.size()) - 1;
        if (n_levels < 3) {
          stack.push_back({std::move(cv), true});
          continue;
        }
        // smallest level whose midpoint lies in the middle 40 % of the vertices
        const double tot = static_cast<double>(bfs_out.size());
        int32_t best_lv = -1;
        for (int32_t L = 1; L + 1 < n_levels; ++L) {
          const double mid = lvl[L] + 0.5 * (lvl[L + 1] - lvl[L]);
          if (mid < 0.3 * tot || mid > 0.7 * tot) continue;
          if (best_lv < 0 ||
              lvl[L + 1] - lvl[L] < lvl[best_lv + 1] - lvl[best_lv] ||
              (lvl[L + 1] - lvl[L] == lvl[best_lv + 1] - lvl[best_lv] &&
               std::abs(mid - 0.5 * tot) <
                   std::abs(lvl[best_lv] +
                            0.5 * (lvl[best_lv + 1] - lvl[best_lv]) -
                            0.5 * tot))) {
            best_lv = L;
          }
        }
        if (best_lv < 0) best_lv = n_levels / 2;
        std::vector<int32_t> A(bfs_out.begin(), bfs_out.begin() + lvl[best_lv]);
        std::vector<int32_t> S(bfs_out.begin() + lvl[best_lv],
                               bfs_out.begin() + lvl[best_lv + 1]);
        std::vector<int32_t> Bv(bfs_out.begin() + lvl[best_lv + 1],
                                bfs_out.end());
        // Trim the separator: a level-set vertex always touches the previous
        // level (side A); one that does not touch side B separates nothing and
        // joins A. On a time-banded KKT graph this turns a separator of a whole
        // stage (states, input and multipliers) into the 4-5 vertices that
        // actually couple the two halves: shorter pivot chains on the critical
        // path of the factorisation, smaller fronts, less fill.
        {
          // mark[v]: 1 = A, 2 = S, 3 = B (within this component)
          for (int32_t v : A) dist[v] = 1;
          for (int32_t v : S) dist[v] = 2;
          for (int32_t v : Bv) dist[v] = 3;
          std::vector<int32_t> keep;
          for (int32_t v : S) {
            bool touches_b = false;
            for (int32_t k = ptr[v]; k < ptr[v + 1] && !touches_b; ++k) {
              const int32_t w = idx[k];
              touches_b = part[w] == cid && dist[w] == 3;
            }
            if (touches_b) {
              keep.push_back(v);
            } else {
              dist[v] = 1;
              A.push_back(v);
            }
          }
          for (int32_t v : bfs_out) dist[v] = -1;
          if (!keep.empty()) S.swap(keep);
          // (an empty trimmed separator means A and B were not connected
          // through S at all; keep the untrimmed one in that odd case)
          else {
            for (int32_t v : S) {
              auto it = std::find(A.begin(), A.end(), v);
              if (it != A.end()) A.erase(it);
            }
          }
        }
        // emitted order: A…, B…, S  ⇒ push S first (stack is LIFO)
        stack.push_back({std::move(S), true});
        stack.push_back({std::move(Bv), false});
        stack.push_back({std::move(A), false});
      }
    }
  }
};

}  // namespace

std::vector<int32_t> order_nested_dissection(const Pattern& lowerK) {
  std::vector<int32_t> ptr, idx;
  symmetric_adjacency(lowerK, ptr, idx);
  int32_t leaf = 20;
  if (const char* e = std::getenv("SLPB_ND_LEAF")) leaf = std::max(2, std::atoi(e));
  NestedDissection nd{ptr, idx, leaf};
  nd.run();
  return nd.order;
}

/// Moves every multiplier (index ≥ n_primal) that would be eliminated before
/// ALL of its neighbours to just behind its first neighbour. Without pivoting a
/// multiplier that comes first meets the pivot −γ (exactly 0 for γ = 0, and
/// element growth 1/γ otherwise); after one neighbour it meets
/// −γ − aᵀ(H + δ)⁻¹a. Minimum-degree orderings have this property on
/// relaxed problems by themselves (the degree-1 slack columns go first);
/// a dissection order needs the nudge.
std::vector<int32_t> defer_leading_multipliers(const Pattern& lowerK,
                                               int32_t n_primal,
                                               const std::vector<int32_t>& perm) {
  const int32_t dim = lowerK.cols;
  std::vector<int32_t> ptr, idx;
  symmetric_adjacency(lowerK, ptr, idx);
  // (1) Primal columns whose only neighbour is one multiplier (the p/n slack
  // columns of a relaxed problem, entry ±1) go first, as a minimum-degree
  // order would put them: no fill, and their multiplier gets a non-zero
  // diagonal whatever the values of its other entries are.
  std::vector<uint8_t> stripped(dim, 0);
  std::vector<int32_t> out;
  out.reserve(dim);
  for (int32_t k = 0; k < dim; ++k) {
    const int32_t v = perm[k];
    if (v >= n_primal) continue;
    int32_t deg = 0, nb = -1;
    for (int32_t q = ptr[v]; q < ptr[v + 1]; ++q) {
      if (idx[q] != v) {
        ++deg;
        nb = idx[q];
      }
    }
    if (deg == 1 && nb >= n_primal) {
      stripped[v] = 1;
      out.push_back(v);
    }
  }
  // (2) Every remaining multiplier that would still come before all of its
  // neighbours moves to just behind the first of them.
  std::vector<int32_t> pos(dim);
  for (int32_t k = 0; k < dim; ++k) pos[perm[k]] = stripped[perm[k]] ? -1 : k;
  std::vector<int32_t> anchor(dim, -1);
  std::vector<int32_t> wait_head(dim, -1), wait_next(dim, -1), wait_tail(dim, -1);
  for (int32_t k = 0; k < dim; ++k) {
    const int32_t d = perm[k];
    if (d < n_primal) continue;
    int32_t first = -1;
    for (int32_t q = ptr[d]; q < ptr[d + 1]; ++q) {
      const int32_t u = idx[q];
      if (u == d) continue;
      if (first < 0 || pos[u] < pos[first]) first = u;
    }
    if (first >= 0 && pos[first] > k) {
      anchor[d] = first;
      // keep the original relative order among the waiters of one anchor
      if (wait_head[first] < 0) {
        wait_head[first] = wait_tail[first] = d;
      } else {
        wait_next[wait_tail[first]] = d;
        wait_tail[first] = d;
      }
    }
  }
  for (int32_t k = 0; k < dim; ++k) {
    const int32_t v = perm[k];
    if (stripped[v] || anchor[v] >= 0) continue;  // already placed / waits
    out.push_back(v);
    for (int32_t d = wait_head[v]; d >= 0; d = wait_next[d]) out.push_back(d);
  }
  return out;
}

bool analyze_kkt(const Pattern& K, int32_t n_primal, int ordering,
                 const int32_t* user_perm, Symbolic& S, std::string& error) {
  S = Symbolic{};
  const int32_t dim = K.cols;
  S.dim = dim;
  std::vector<int32_t> perm;
  switch (ordering) {
    case SLPB_ORDER_NESTED_DISSECTION:
      perm = defer_leading_multipliers(K, n_primal, order_nested_dissection(K));
      break;
    case SLPB_ORDER_AMD:
      // the reference's own order (Eigen::SimplicialLDLT's default AMDOrdering,
      // sparse_regularized_ldlt.hpp:183), taken as is: no multiplier deferral
      perm = order_amd(K);
      break;
    case SLPB_ORDER_NATURAL:
      perm.resize(dim);
      std::iota(perm.begin(), perm.end(), 0);
      break;
    case SLPB_ORDER_CUSTOM:
      if (!user_perm) {
        error = "slpb_analyze: SLPB_ORDER_CUSTOM needs a permutation";
        return false;
      }
      perm.assign(user_perm, user_perm + dim);
      break;
    default:
      error = "slpb_analyze: this ordering is not implemented; pass a "
              "permutation with SLPB_ORDER_CUSTOM";
      return false;
  }
  {
    std::vector<uint8_t> seen(dim, 0);
    if (static_cast<int32_t>(perm.size()) != dim) {
      error = "slpb_analyze: permutation has the wrong length";
      return false;
    }
    for (int32_t p : perm) {
      if (p < 0 || p >= dim || seen[p]) {
        error = "slpb_analyze: not a permutation";
        return false;
      }
      seen[p] = 1;
    }
  }

  // permuted matrix as "row lists of the upper triangle": for column k, the
  // rows i < k with B(i,k) ≠ 0, B = K(perm, perm)
  auto build_upper = [&](const std::vector<int32_t>& p,
                         std::vector<int32_t>& up, std::vector<int32_t>& ui) {
    std::vector<int32_t> ip(dim);
    for (int32_t k = 0; k < dim; ++k) ip[p[k]] = k;
    up.assign(dim + 1, 0);
    for (int32_t c = 0; c < dim; ++c) {
      for (int32_t k = K.colptr[c]; k < K.colptr[c + 1]; ++k) {
        const int32_t a = ip[K.rowidx[k]], b = ip[c];
        if (a != b) ++up[std::max(a, b) + 1];
      }
    }
    for (int32_t i = 0; i < dim; ++i) up[i + 1] += up[i];
    ui.assign(up[dim], 0);
    std::vector<int32_t> nxt(up.begin(), up.end() - 1);
    for (int32_t c = 0; c < dim; ++c) {
      for (int32_t k = K.colptr[c]; k < K.colptr[c + 1]; ++k) {
        const int32_t a = ip[K.rowidx[k]], b = ip[c];
        if (a != b) ui[nxt[std::max(a, b)]++] = std::min(a, b);
      }
    }
  };
  auto etree = [&](const std::vector<int32_t>& up,
                   const std::vector<int32_t>& ui, std::vector<int32_t>& par) {
    par.assign(dim, -1);
    std::vector<int32_t> anc(dim, -1);
    for (int32_t k = 0; k < dim; ++k) {
      for (int32_t q = up[k]; q < up[k + 1]; ++q) {
        int32_t i = ui[q];
        while (i != -1 && i < k) {
          const int32_t nx = anc[i];
          anc[i] = k;
          if (nx == -1) par[i] = k;
          i = nx;
        }
      }
    }
  };

  std::vector<int32_t> up, ui, par;
  build_upper(perm, up, ui);
  etree(up, ui, par);
  // postorder the elimination tree so that every subtree is contiguous, then
  // fold the postorder into the permutation
  {
    std::vector<int32_t> head(dim, -1), next(dim, -1), post;
    post.reserve(dim);
    for (int32_t j = dim - 1; j >= 0; --j) {
      if (par[j] >= 0) {
        next[j] = head[par[j]];
        head[par[j]] = j;
      }
    }
    std::vector<int32_t> stack;
    for (int32_t r = 0; r < dim; ++r) {
      if (par[r] != -1) continue;
      stack.push_back(r);
      while (!stack.empty()) {
        const int32_t p = stack.back();
        const int32_t c = head[p];
        if (c == -1) {
          stack.pop_back();
          post.push_back(p);
        } else {
          head[p] = next[c];
          stack.push_back(c);
        }
      }
    }
    std::vector<int32_t> perm2(dim);
    for (int32_t k = 0; k < dim; ++k) perm2[k] = perm[post[k]];
    perm.swap(perm2);
    build_upper(perm, up, ui);
    etree(up, ui, par);
  }
  S.perm = perm;
  S.iperm.assign(dim, 0);
  for (int32_t k = 0; k < dim; ++k) S.iperm[perm[k]] = k;
  S.parent = par;

  // column structures of L (rows strictly below the diagonal, ascending)
  std::vector<std::vector<int32_t>> Ls(dim);
  {
    // lower entries of B by column: B(i,k), i>k ↔ upper entry (k, i) in col i
    std::vector<std::vector<int32_t>> lower(dim);
    for (int32_t k = 0; k < dim; ++k) {
      for (int32_t q = up[k]; q < up[k + 1]; ++q) lower[ui[q]].push_back(k);
    }
    std::vector<std::vector<int32_t>> kids(dim);
    for (int32_t j = 0; j < dim; ++j) {
      if (par[j] >= 0) kids[par[j]].push_back(j);
    }
    std::vector<int32_t> mark(dim, -1);
    for (int32_t j = 0; j < dim; ++j) {
      auto& col = Ls[j];
      mark[j] = j;
      for (int32_t r : lower[j]) {
        if (mark[r] != j) {
          mark[r] = j;
          col.push_back(r);
        }
      }
      for (int32_t c : kids[j]) {
        for (int32_t r : Ls[c]) {
          if (r != j && mark[r] != j) {
            mark[r] = j;
            col.push_back(r);
          }
        }
      }
      std::sort(col.begin(), col.end());
      S.nnz_l += static_cast<int64_t>(col.size());
    }
    std::vector<int32_t> depth(dim, 1);
    for (int32_t j = 0; j < dim; ++j) {
      if (par[j] >= 0) depth[par[j]] = std::max(depth[par[j]], depth[j] + 1);
      S.etree_height = std::max(S.etree_height, depth[j]);
    }
  }

  // fundamental supernodes: j+1 joins j when it is j's parent and its
  // structure is j's minus {j+1}
  std::vector<int32_t> first;  // first column of each supernode
  for (int32_t j = 0; j < dim; ++j) {
    const bool join = j > 0 && par[j - 1] == j &&
                      Ls[j].size() + 1 == Ls[j - 1].size();
    if (!join) first.push_back(j);
  }
  first.push_back(dim);
  int32_t ns = static_cast<int32_t>(first.size()) - 1;

  // relaxed amalgamation of a supernode with its LAST child (the one that
  // ends right before it), bottom-up; thresholds in the spirit of CHOLMOD's
  // nrelax/zrelax, tuned towards fronts of at most ~32 rows per warp
  {
    std::vector<int32_t> sfirst(first.begin(), first.end() - 1), slast(ns);
    for (int32_t s = 0; s < ns; ++s) slast[s] = first[s + 1] - 1;
    std::vector<int64_t> zeros(ns, 0);  // explicit zeros already in the panel
    std::vector<int32_t> out_first;
    std::vector<int64_t> out_zeros;
    for (int32_t s = 0; s < ns; ++s) {
      int32_t f = sfirst[s];
      int64_t z = zeros[s];
      const int32_t l = slast[s];
      // merge candidates are the already-emitted supernodes at the tail
      while (!out_first.empty()) {
        const int32_t cf = out_first.back();
        const int32_t cl = f - 1;               // child's last column
        if (par[cl] < f || par[cl] > l) break;  // not a child of this one
        const int32_t nc = cl - cf + 1;         // child columns
        const int32_t np = l - f + 1;           // parent columns
        const int64_t below_p = static_cast<int64_t>(Ls[l].size());
        const int64_t dim_p = np + below_p;     // parent front order
        const int64_t below_c = static_cast<int64_t>(Ls[cl].size());
        // merged panel: (nc + dim_p) rows × (nc + np) cols lower trapezoid
        const int64_t new_zeros =
            static_cast<int64_t>(nc) * (dim_p - below_c);  // child cols padded
        const int64_t tot_zeros = z + out_zeros.back() + new_zeros;
        const int64_t ncols = nc + np;
        const int64_t dim_m = nc + dim_p;
        const double entries =
            static_cast<double>(ncols) * dim_m -
            static_cast<double>(ncols) * (ncols - 1) / 2.0;
        const double frac = tot_zeros / entries;
        const bool merge = dim_m <= 8 || (dim_m <= 20 && frac < 0.6) ||
                           (dim_m <= 32 && frac < 0.35) ||
                           (dim_m <= 64 && frac < 0.15) || frac < 0.05;
        if (!merge) break;
        f = cf;
        z = tot_zeros;
        out_first.pop_back();
        out_zeros.pop_back();
      }
      out_first.push_back(f);
      out_zeros.push_back(z);
    }
    first = out_first;
    first.push_back(dim);
    ns = static_cast<int32_t>(first.size()) - 1;
  }
  S.n_super = ns;
  S.super_first = first;

  std::vector<int32_t> super_of(dim);
  for (int32_t s = 0; s < ns; ++s) {
    for (int32_t j = first[s]; j < first[s + 1]; ++j) super_of[j] = s;
  }
  // front rows: own columns, then the structure below the last own column
  // (after amalgamation the structure is the union over the own columns)
  S.rows_ptr.assign(ns + 1, 0);
  S.front_dim.assign(ns, 0);
  S.super_parent.assign(ns, -1);
  {
    std::vector<int32_t> mark(dim, -1);
    std::vector<int32_t> below;
    for (int32_t s = 0; s < ns; ++s) {
      const int32_t f = first[s], l = first[s + 1] - 1;
      below.clear();
      for (int32_t j = f; j <= l; ++j) {
        for (int32_t r : Ls[j]) {
          if (r > l && mark[r] != s) {
            mark[r] = s;
            below.push_back(r);
          }
        }
      }
      std::sort(below.begin(), below.end());
      for (int32_t j = f; j <= l; ++j) S.rows_idx.push_back(j);
      S.rows_idx.insert(S.rows_idx.end(), below.begin(), below.end());
      S.rows_ptr[s + 1] = static_cast<int64_t>(S.rows_idx.size());
      S.front_dim[s] = (l - f + 1) + static_cast<int32_t>(below.size());
      S.max_front = std::max(S.max_front, S.front_dim[s]);
      if (par[l] >= 0) S.super_parent[s] = super_of[par[l]];
    }
  }
  // children, levels
  S.child_ptr.assign(ns + 1, 0);
  for (int32_t s = 0; s < ns; ++s) {
    if (S.super_parent[s] >= 0) ++S.child_ptr[S.super_parent[s] + 1];
  }
  for (int32_t s = 0; s < ns; ++s) S.child_ptr[s + 1] += S.child_ptr[s];
  S.child_idx.assign(S.child_ptr[ns], 0);
  {
    std::vector<int64_t> nxt(S.child_ptr.begin(), S.child_ptr.end() - 1);
    for (int32_t s = 0; s < ns; ++s) {
      if (S.super_parent[s] >= 0) S.child_idx[nxt[S.super_parent[s]]++] = s;
    }
  }
  S.super_level.assign(ns, 0);
  for (int32_t s = 0; s < ns; ++s) {
    const int32_t p = S.super_parent[s];
    if (p >= 0) S.super_level[p] = std::max(S.super_level[p], S.super_level[s] + 1);
  }
  S.n_levels = 0;
  for (int32_t s = 0; s < ns; ++s) S.n_levels = std::max(S.n_levels, S.super_level[s] + 1);
  S.level_ptr.assign(S.n_levels + 1, 0);
  for (int32_t s = 0; s < ns; ++s) ++S.level_ptr[S.super_level[s] + 1];
  for (int32_t L = 0; L < S.n_levels; ++L) S.level_ptr[L + 1] += S.level_ptr[L];
  S.level_supers.assign(ns, 0);
  {
    std::vector<int32_t> nxt(S.level_ptr.begin(), S.level_ptr.end() - 1);
    for (int32_t s = 0; s < ns; ++s) S.level_supers[nxt[S.super_level[s]]++] = s;
  }
  // storage + relative indices
  S.panel_ptr.assign(ns + 1, 0);
  S.update_ptr.assign(ns + 1, 0);
  S.rel_ptr.assign(ns + 1, 0);
  for (int32_t s = 0; s < ns; ++s) {
    const int64_t F = S.front_dim[s];
    const int64_t np = first[s + 1] - first[s];
    const int64_t m = F - np;
    // packed: the lower trapezoid of the panel, the lower triangle of the update
    // matrix (column j of a triangle of order n: entries i ≥ j at tri_col(j, n) + i)
    S.panel_ptr[s + 1] = S.panel_ptr[s] + F * np - (np * (np - 1)) / 2;
    S.update_ptr[s + 1] = S.update_ptr[s] + (m * (m + 1)) / 2;
    S.rel_ptr[s + 1] = S.rel_ptr[s] + m;
  }
  S.panel_size = S.panel_ptr[ns];
  S.update_size = S.update_ptr[ns];
  S.rel_idx.assign(S.rel_ptr[ns], 0);
  {
    std::vector<int32_t> where(dim, -1);
    for (int32_t p = 0; p < ns; ++p) {
      const int64_t rb = S.rows_ptr[p], re = S.rows_ptr[p + 1];
      for (int64_t k = rb; k < re; ++k) where[S.rows_idx[k]] = int32_t(k - rb);
      for (int64_t ck = S.child_ptr[p]; ck < S.child_ptr[p + 1]; ++ck) {
        const int32_t c = S.child_idx[ck];
        const int64_t npc = first[c + 1] - first[c];
        const int64_t cb = S.rows_ptr[c] + npc, ce = S.rows_ptr[c + 1];
        for (int64_t k = cb; k < ce; ++k) {
          const int32_t w = where[S.rows_idx[k]];
          if (w < 0) {
            error = "internal: child update row missing from parent front";
            return false;
          }
          S.rel_idx[S.rel_ptr[c] + (k - cb)] = w;
        }
      }
      for (int64_t k = rb; k < re; ++k) where[S.rows_idx[k]] = -1;
    }
  }
  // scatter of K entries into fronts
  {
    std::vector<std::vector<std::pair<int32_t, int32_t>>> per(ns);
    std::vector<int32_t> where(dim, -1);
    // group K entries by owning supernode first (owner = supernode of the
    // smaller permuted index)
    for (int32_t c = 0; c < dim; ++c) {
      for (int32_t k = K.colptr[c]; k < K.colptr[c + 1]; ++k) {
        const int32_t a = S.iperm[K.rowidx[k]], b = S.iperm[c];
        per[super_of[std::min(a, b)]].emplace_back(k, 0);
      }
    }
    S.asm_ptr.assign(ns + 1, 0);
    for (int32_t s = 0; s < ns; ++s) {
      const int64_t rb = S.rows_ptr[s], re = S.rows_ptr[s + 1];
      const int32_t F = S.front_dim[s];
      for (int64_t k = rb; k < re; ++k) where[S.rows_idx[k]] = int32_t(k - rb);
      for (auto& [ke, dst] : per[s]) {
        // recover (row, col) of K entry ke
        const int32_t c = static_cast<int32_t>(
            std::upper_bound(K.colptr.begin(), K.colptr.end(), ke) -
            K.colptr.begin() - 1);
        const int32_t a = S.iperm[K.rowidx[ke]], b = S.iperm[c];
        const int32_t lr = where[std::max(a, b)], lc = where[std::min(a, b)];
        if (lr < 0 || lc < 0) {
          error = "internal: KKT entry outside its front";
          return false;
        }
        S.asm_src.push_back(ke);
        S.asm_dst.push_back(lr + lc * F);
      }
      S.asm_ptr[s + 1] = static_cast<int64_t>(S.asm_src.size());
      for (int64_t k = rb; k < re; ++k) where[S.rows_idx[k]] = -1;
    }
  }
  S.col_is_primal.assign(dim, 0);
  for (int32_t k = 0; k < dim; ++k) S.col_is_primal[k] = perm[k] < n_primal;
  return true;
}

void build_tree_shard(const Symbolic& Y, int32_t world, TreeShard& out) {
  out = TreeShard{};
  out.world = std::max(world, 1);
  const int32_t ns = Y.n_super;
  out.owner.assign(ns, -1);
  out.top_fcount_init.assign(ns, 0);
  out.rank_order.assign(out.world, {});
  out.rank_roots.assign(out.world, {});
  out.rank_work.assign(out.world, 0.0);
  // work of every front and of every subtree (children precede parents)
  std::vector<double> work(ns), subtree(ns);
  for (int32_t s = 0; s < ns; ++s) {
    const double F = Y.front_dim[s];
    const double np = Y.super_first[s + 1] - Y.super_first[s];
    work[s] = np * F * F + 64.0;
    subtree[s] = work[s];
  }
  double total = 0.0;
  for (int32_t s = 0; s < ns; ++s) {
    if (Y.super_parent[s] >= 0) {
      subtree[Y.super_parent[s]] += subtree[s];
    } else {
      total += subtree[s];
    }
  }
  // frontier of subtree roots: start at the roots of the forest and keep
  // opening the heaviest one (it joins the top, its children the frontier)
  // until there are enough pieces to balance and none of them dominates
  std::vector<int32_t> frontier;
  std::vector<uint8_t> is_top(ns, 0);
  for (int32_t s = 0; s < ns; ++s) {
    if (Y.super_parent[s] < 0) frontier.push_back(s);
  }
  if (out.world > 1) {
    const size_t want = static_cast<size_t>(8 * out.world);
    const double fine = total / (16.0 * out.world);
    for (;;) {
      int32_t best = -1;
      for (size_t k = 0; k < frontier.size(); ++k) {
        const int32_t s = frontier[k];
        if (Y.child_ptr[s + 1] == Y.child_ptr[s]) continue;  // a leaf stays
        if (best < 0 || subtree[s] > subtree[frontier[best]]) {
          best = static_cast<int32_t>(k);
        }
      }
      if (best < 0) break;
      const int32_t s = frontier[best];
      if (frontier.size() >= want && subtree[s] <= fine) break;
      is_top[s] = 1;
      frontier.erase(frontier.begin() + best);
      for (int64_t c = Y.child_ptr[s]; c < Y.child_ptr[s + 1]; ++c) {
        frontier.push_back(Y.child_idx[c]);
      }
    }
  }
  // heaviest subtree first, each to the least loaded rank
  std::sort(frontier.begin(), frontier.end(), [&](int32_t a, int32_t b) {
    return subtree[a] != subtree[b] ? subtree[a] > subtree[b] : a < b;
  });
  std::vector<int32_t> root_owner(ns, -1);
  for (int32_t r : frontier) {
    int32_t rank = 0;
    for (int32_t q = 1; q < out.world; ++q) {
      if (out.rank_work[q] < out.rank_work[rank]) rank = q;
    }
    root_owner[r] = rank;
    out.rank_work[rank] += subtree[r];
    out.rank_roots[rank].push_back(r);
  }
  // owners flow down from the subtree roots (parents have larger indices)
  for (int32_t s = ns - 1; s >= 0; --s) {
    if (is_top[s]) {
      out.owner[s] = -1;
      out.top_work += work[s];
    } else if (root_owner[s] >= 0) {
      out.owner[s] = root_owner[s];
    } else {
      out.owner[s] = out.owner[Y.super_parent[s]];
    }
  }
  for (int32_t k = 0; k < ns; ++k) {
    const int32_t s = Y.level_supers[k];  // ascending level
    if (out.owner[s] < 0) {
      out.top_order.push_back(s);
    } else {
      out.rank_order[out.owner[s]].push_back(s);
    }
  }
  for (int32_t s = 0; s < ns; ++s) {
    const int32_t p = Y.super_parent[s];
    if (p >= 0 && out.owner[p] < 0 && out.owner[s] >= 0) ++out.top_fcount_init[p];
  }
  for (auto& roots : out.rank_roots) std::sort(roots.begin(), roots.end());
}

}  // namespace slpb
