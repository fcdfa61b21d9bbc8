// Approximate minimum degree ordering (SLPB_ORDER_AMD).
//
// The reference factors its KKT systems with Eigen::SimplicialLDLT, whose
// default ordering is AMDOrdering<int> applied to the full symmetric pattern
// (solver/util/sparse_regularized_ldlt.hpp:66-74,183). Eigen is not in this
// image; this file restates the published algorithm it implements — Amestoy,
// Davis & Duff, "An approximate minimum degree ordering algorithm" (SIMAX 1996;
// ACM TOMS 837, 2004) in its quotient-graph form with element absorption,
// aggressive absorption, mass elimination, hashed supervariable detection,
// deferred dense rows and an assembly-tree postorder — so that the device can
// be driven in the reference's own elimination order (at the price of an
// elimination tree of height O(N) on transcription problems; the default
// ordering of the device path stays nested dissection).
//
// The quotient graph lives in one integer pool. A live *variable* i owns the
// slice [start[i], start[i] + len[i]): first elen[i] adjacent elements, then
// adjacent variables. A live *element* e owns the slice of its member
// variables. Dead objects store the (negated) index of what absorbed them in
// start[]. Supervariables carry their multiplicity in nv[].
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

#include "internal.hpp"

namespace slpb {
namespace {

constexpr int32_t kNone = -1;
inline int32_t dead_ref(int32_t i) { return -i - 2; }  // involution

class QuotientGraph {
 public:
  QuotientGraph(int32_t n, const std::vector<int32_t>& ptr,
                const std::vector<int32_t>& idx)
      : n_{n}, start_(ptr.begin(), ptr.end()), len_(n + 1, 0), nv_(n + 1, 1),
        elen_(n + 1, 0), degree_(n + 1, 0), stamp_(n + 1, 1),
        next_(n + 1, kNone), prev_(n + 1, kNone), bucket_(n + 1, kNone),
        hash_head_(n + 1, kNone) {
    used_ = ptr[n];
    capacity_ = used_ + used_ / 5 + 2 * n;
    pool_.assign(capacity_, 0);
    std::copy(idx.begin(), idx.begin() + used_, pool_.begin());
    for (int32_t i = 0; i < n; ++i) {
      len_[i] = ptr[i + 1] - ptr[i];
      degree_[i] = len_[i];
    }
    dense_ = std::min(n - 2, std::max<int32_t>(
                                 16, static_cast<int32_t>(10 * std::sqrt(double(n)))));
    mark_ = reset_stamps(0);
  }

  std::vector<int32_t> run() {
    seed_degree_lists();
    while (eliminated_ < n_) {
      const int32_t k = take_min_degree_pivot();
      const int32_t n_elems = elen_[k];
      int32_t nvk = nv_[k];
      eliminated_ += nvk;
      if (n_elems > 0 && used_ + mindeg_ >= capacity_) compact();
      int32_t dk = 0;
      const auto [lo, hi] = form_element(k, n_elems, nvk, dk);
      mark_ = reset_stamps(mark_);
      stamp_external_sizes(lo, hi);
      update_members(k, lo, hi, dk, nvk);
      degree_[k] = dk;
      lemax_ = std::max(lemax_, dk);
      mark_ = reset_stamps(mark_ + lemax_);
      merge_indistinguishable(lo, hi);
      finish_element(k, lo, hi, dk, nvk, n_elems);
    }
    return postorder();
  }

 private:
  int32_t reset_stamps(int32_t mark) {
    if (mark < 2 || mark + lemax_ < 0) {
      for (int32_t i = 0; i < n_; ++i) {
        if (stamp_[i] != 0) stamp_[i] = 1;
      }
      mark = 2;
    }
    return mark;
  }

  void list_push(int32_t d, int32_t i) {
    if (bucket_[d] != kNone) prev_[bucket_[d]] = i;
    next_[i] = bucket_[d];
    bucket_[d] = i;
  }
  void list_remove(int32_t i) {
    if (next_[i] != kNone) prev_[next_[i]] = prev_[i];
    if (prev_[i] != kNone) {
      next_[prev_[i]] = next_[i];
    } else {
      bucket_[degree_[i]] = next_[i];
    }
  }

  /// A row holding only its diagonal is eliminated at once; a dense row, or
  /// one without a structural diagonal, is absorbed by the artificial root n
  /// and ordered last; everything else enters the degree lists.
  void seed_degree_lists() {
    for (int32_t i = 0; i < n_; ++i) {
      bool diag = false;
      for (int32_t p = start_[i]; p < start_[i] + len_[i] && !diag; ++p) {
        diag = pool_[p] == i;
      }
      const int32_t d = degree_[i];
      if (d == 1 && diag) {
        elen_[i] = -2;
        ++eliminated_;
        start_[i] = kNone;
        stamp_[i] = 0;
      } else if (d > dense_ || !diag) {
        nv_[i] = 0;
        elen_[i] = -1;
        ++eliminated_;
        start_[i] = dead_ref(n_);
        ++nv_[n_];
      } else {
        list_push(d, i);
      }
    }
    elen_[n_] = -2;
    start_[n_] = kNone;
    stamp_[n_] = 0;
  }

  int32_t take_min_degree_pivot() {
    int32_t k = kNone;
    while (mindeg_ < n_ && (k = bucket_[mindeg_]) == kNone) ++mindeg_;
    if (next_[k] != kNone) prev_[next_[k]] = kNone;
    bucket_[mindeg_] = next_[k];
    return k;
  }

  /// Garbage collection of the pool: live slices are packed to the front in
  /// their current order.
  void compact() {
    for (int32_t j = 0; j < n_; ++j) {
      const int32_t p = start_[j];
      if (p >= 0) {
        start_[j] = pool_[p];
        pool_[p] = dead_ref(j);
      }
    }
    int32_t q = 0;
    for (int32_t p = 0; p < used_;) {
      const int32_t j = dead_ref(pool_[p++]);
      if (j >= 0) {
        pool_[q] = start_[j];
        start_[j] = q++;
        for (int32_t t = 0; t < len_[j] - 1; ++t) pool_[q++] = pool_[p++];
      }
    }
    used_ = q;
  }

  /// Builds the member list of the new element k: the live variables adjacent
  /// to k directly or through one of k's elements (which k absorbs). Members
  /// leave the degree lists and get nv negated as an "in Lk" flag.
  std::pair<int32_t, int32_t> form_element(int32_t k, int32_t n_elems,
                                           int32_t nvk, int32_t& dk) {
    nv_[k] = -nvk;
    int32_t p = start_[k];
    const int32_t lo = n_elems == 0 ? p : used_;
    int32_t hi = lo;
    for (int32_t t = 1; t <= n_elems + 1; ++t) {
      int32_t e, src, count;
      if (t > n_elems) {
        e = k;
        src = p;
        count = len_[k] - n_elems;
      } else {
        e = pool_[p++];
        src = start_[e];
        count = len_[e];
      }
      for (int32_t c = 0; c < count; ++c) {
        const int32_t i = pool_[src++];
        const int32_t nvi = nv_[i];
        if (nvi <= 0) continue;
        dk += nvi;
        nv_[i] = -nvi;
        pool_[hi++] = i;
        list_remove(i);
      }
      if (e != k) {
        start_[e] = dead_ref(k);
        stamp_[e] = 0;
      }
    }
    if (n_elems != 0) used_ = hi;
    degree_[k] = dk;
    start_[k] = lo;
    len_[k] = hi - lo;
    elen_[k] = -2;
    return {lo, hi};
  }

  /// stamp[e] − mark = |Le \ Lk| for every element e next to a member of Lk.
  void stamp_external_sizes(int32_t lo, int32_t hi) {
    for (int32_t pk = lo; pk < hi; ++pk) {
      const int32_t i = pool_[pk];
      const int32_t ne = elen_[i];
      if (ne <= 0) continue;
      const int32_t nvi = -nv_[i];
      const int32_t fresh = mark_ - nvi;
      for (int32_t q = start_[i]; q < start_[i] + ne; ++q) {
        const int32_t e = pool_[q];
        if (stamp_[e] >= mark_) {
          stamp_[e] -= nvi;
        } else if (stamp_[e] != 0) {
          stamp_[e] = degree_[e] + fresh;
        }
      }
    }
  }

  /// Approximate degrees of the members, aggressive element absorption, mass
  /// elimination, and hashing for the supervariable search.
  void update_members(int32_t k, int32_t lo, int32_t hi, int32_t& dk,
                      int32_t& nvk) {
    for (int32_t pk = lo; pk < hi; ++pk) {
      const int32_t i = pool_[pk];
      const int32_t p1 = start_[i];
      const int32_t p2 = p1 + elen_[i] - 1;
      int32_t out = p1, d = 0;
      uint32_t h = 0;
      for (int32_t q = p1; q <= p2; ++q) {
        const int32_t e = pool_[q];
        if (stamp_[e] == 0) continue;
        const int32_t ext = stamp_[e] - mark_;
        if (ext > 0) {
          d += ext;
          pool_[out++] = e;
          h += static_cast<uint32_t>(e);
        } else {
          start_[e] = dead_ref(k);  // Le ⊆ Lk: absorbed
          stamp_[e] = 0;
        }
      }
      elen_[i] = out - p1 + 1;
      const int32_t first_var = out;
      const int32_t end = p1 + len_[i];
      for (int32_t q = p2 + 1; q < end; ++q) {
        const int32_t j = pool_[q];
        const int32_t nvj = nv_[j];
        if (nvj <= 0) continue;
        d += nvj;
        pool_[out++] = j;
        h += static_cast<uint32_t>(j);
      }
      if (d == 0) {
        start_[i] = dead_ref(k);  // indistinguishable from the pivot
        const int32_t nvi = -nv_[i];
        dk -= nvi;
        nvk += nvi;
        eliminated_ += nvi;
        nv_[i] = 0;
        elen_[i] = -1;
      } else {
        degree_[i] = std::min(degree_[i], d);
        pool_[out] = pool_[first_var];
        pool_[first_var] = pool_[p1];
        pool_[p1] = k;  // k becomes i's first element
        len_[i] = out - p1 + 1;
        h %= static_cast<uint32_t>(n_);
        next_[i] = hash_head_[h];
        hash_head_[h] = i;
        prev_[i] = static_cast<int32_t>(h);
      }
    }
  }

  void merge_indistinguishable(int32_t lo, int32_t hi) {
    for (int32_t pk = lo; pk < hi; ++pk) {
      int32_t i = pool_[pk];
      if (nv_[i] >= 0) continue;
      const int32_t h = prev_[i];
      i = hash_head_[h];
      hash_head_[h] = kNone;
      for (; i != kNone && next_[i] != kNone; i = next_[i], ++mark_) {
        const int32_t ln = len_[i], ne = elen_[i];
        for (int32_t q = start_[i] + 1; q < start_[i] + ln; ++q) {
          stamp_[pool_[q]] = mark_;
        }
        int32_t tail = i;
        for (int32_t j = next_[i]; j != kNone;) {
          bool same = len_[j] == ln && elen_[j] == ne;
          for (int32_t q = start_[j] + 1; same && q < start_[j] + ln; ++q) {
            same = stamp_[pool_[q]] == mark_;
          }
          if (same) {
            start_[j] = dead_ref(i);
            nv_[i] += nv_[j];
            nv_[j] = 0;
            elen_[j] = -1;
            j = next_[j];
            next_[tail] = j;
          } else {
            tail = j;
            j = next_[j];
          }
        }
      }
    }
  }

  void finish_element(int32_t k, int32_t lo, int32_t hi, int32_t dk,
                      int32_t nvk, int32_t n_elems) {
    int32_t out = lo;
    for (int32_t pk = lo; pk < hi; ++pk) {
      const int32_t i = pool_[pk];
      const int32_t nvi = -nv_[i];
      if (nvi <= 0) continue;
      nv_[i] = nvi;
      int32_t d = degree_[i] + dk - nvi;
      d = std::min(d, n_ - eliminated_ - nvi);
      prev_[i] = kNone;
      list_push(d, i);
      mindeg_ = std::min(mindeg_, d);
      degree_[i] = d;
      pool_[out++] = i;
    }
    nv_[k] = nvk;
    len_[k] = out - lo;
    if (len_[k] == 0) {
      start_[k] = kNone;
      stamp_[k] = 0;
    }
    if (n_elems != 0) used_ = out;
  }

  /// Depth-first postorder of the assembly tree (absorbed variables hang below
  /// the element that took them; elements below the one that absorbed them).
  std::vector<int32_t> postorder() {
    std::vector<int32_t>& parent = start_;
    for (int32_t i = 0; i < n_; ++i) parent[i] = dead_ref(parent[i]);
    std::vector<int32_t> child(n_ + 1, kNone), sibling(n_ + 1, kNone);
    for (int32_t j = n_; j >= 0; --j) {
      if (nv_[j] > 0) continue;
      sibling[j] = child[parent[j]];
      child[parent[j]] = j;
    }
    for (int32_t e = n_; e >= 0; --e) {
      if (nv_[e] <= 0) continue;
      if (parent[e] != kNone) {
        sibling[e] = child[parent[e]];
        child[parent[e]] = e;
      }
    }
    std::vector<int32_t> post;
    post.reserve(n_ + 1);
    std::vector<int32_t> stack;
    for (int32_t r = 0; r <= n_; ++r) {
      if (parent[r] != kNone) continue;
      stack.push_back(r);
      while (!stack.empty()) {
        const int32_t p = stack.back();
        const int32_t c = child[p];
        if (c == kNone) {
          stack.pop_back();
          post.push_back(p);
        } else {
          child[p] = sibling[c];
          stack.push_back(c);
        }
      }
    }
    post.resize(n_);
    return post;
  }

  int32_t n_;
  std::vector<int32_t> pool_, start_, len_, nv_, elen_, degree_, stamp_, next_,
      prev_, bucket_, hash_head_;
  int32_t used_ = 0, capacity_ = 0, dense_ = 0;
  int32_t mark_ = 0, lemax_ = 0, mindeg_ = 0, eliminated_ = 0;
};

}  // namespace

std::vector<int32_t> order_amd(const Pattern& lowerK) {
  const int32_t n = lowerK.cols;
  if (n == 0) return {};
  // full symmetric pattern, diagonal kept, every column sorted by row
  std::vector<int32_t> ptr(n + 1, 0);
  for (int32_t c = 0; c < n; ++c) {
    for (int32_t k = lowerK.colptr[c]; k < lowerK.colptr[c + 1]; ++k) {
      const int32_t r = lowerK.rowidx[k];
      ++ptr[c + 1];
      if (r != c) ++ptr[r + 1];
    }
  }
  for (int32_t c = 0; c < n; ++c) ptr[c + 1] += ptr[c];
  std::vector<int32_t> idx(ptr[n]), nxt(ptr.begin(), ptr.end() - 1);
  for (int32_t c = 0; c < n; ++c) {  // rows above the diagonal first
    for (int32_t k = lowerK.colptr[c]; k < lowerK.colptr[c + 1]; ++k) {
      const int32_t r = lowerK.rowidx[k];
      if (r != c) idx[nxt[r]++] = c;
    }
  }
  for (int32_t c = 0; c < n; ++c) {
    for (int32_t k = lowerK.colptr[c]; k < lowerK.colptr[c + 1]; ++k) {
      idx[nxt[c]++] = lowerK.rowidx[k];
    }
  }
  return QuotientGraph{n, ptr, idx}.run();
}

}  // namespace slpb
