// Host-side compiler: uploaded expression tape + row descriptors → cluster
// programs for the autodiff sweep kernels (see internal.hpp for the formats).
//
// What it preserves from the reference (include/sleipnir/autodiff):
//   * the per-row parent→child order of expression_graph.hpp:28-78 — adjoint
//     contributions into a node are pulled in exactly the order in which
//     append_triplets (:119-145) would have pushed them, so sums round alike;
//   * "every wrt leaf of a row emits a triplet, even a zero one" (:147-152),
//     which keeps the sparsity patterns static;
//   * LINEAR rows are constants (jacobian.hpp:84-89): they never reach the
//     device programs, only the constant part of the stage arrays.
// What it adds: rows that share interior nodes are fused into one cluster and
// evaluated once (the reference re-walks shared sub-graphs per row,
// jacobian.hpp:139-141); long root sums (Σ_k u_k² costs) are split into
// independent terms so a 5000-term chain does not serialise on one warp.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <numeric>
#include <queue>
#include <thread>
#include <unordered_map>

#include "internal.hpp"

namespace slpb {

int64_t Pattern::find(int32_t r, int32_t c) const {
  const int32_t* b = rowidx.data() + colptr[c];
  const int32_t* e = rowidx.data() + colptr[c + 1];
  const int32_t* it = std::lower_bound(b, e, r);
  if (it == e || *it != r) return -1;
  return it - rowidx.data();
}

namespace {

/// SLPB_COMPILE_TIMING=1 prints where the host compiler spends its time.
struct StageTimer {
  bool on = std::getenv("SLPB_COMPILE_TIMING") != nullptr;
  std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
  void lap(const char* what) {
    if (!on) return;
    const auto now = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[slpb compile] %-28s %8.1f ms\n", what,
                 std::chrono::duration<double, std::milli>(now - t).count());
    t = now;
  }
};

constexpr int kSplitThreshold = 16;  // root sums shorter than this stay whole
// Sub-rows are fused into one cluster while the sum of their node lists stays
// below this (cart-pole's per-step cluster: 1 642, g-fold's: 400).
constexpr int64_t kMaxClusterVisits = 4000;

struct SubRow {
  int out = 0;          // slpb_output
  int32_t row = 0;
  int8_t seed = 1;
  bool is_value = false;
  std::vector<int32_t> nodes;                       // parent→child order
  std::vector<uint8_t> flags;                       // per node: kActive|kValue
  std::vector<std::pair<int32_t, int32_t>> outs;    // (stage slot, node)
  int32_t value_stage = -1;                         // value rows
};

constexpr uint8_t kActive = 1;  // the node's adjoint reaches an output leaf
constexpr uint8_t kValue = 2;   // the node's value is read by some partial

/// Partials that the forward sweep has already computed as a value are read
/// from that node's slot (emit_program); SLPB_NO_VALUE_REUSE=1 evaluates them
/// again (development switch).
bool value_reuse_enabled() {
  static const bool on = std::getenv("SLPB_NO_VALUE_REUSE") == nullptr;
  return on;
}

/// Which values the partial of `op` w.r.t. `side` reads (bit 0: lhs value,
/// bit 1: rhs value, bit 2: the node's own value); see ad_op_grad in
/// ad_core.hpp.
uint8_t grad_value_needs(uint8_t op, int side) {
  switch (op) {
    // d exp(x) = exp(x): the node's own value (x stays marked: the forward
    // sweep needs it to evaluate the node)
    case SLPB_OP_EXP: return value_reuse_enabled() ? 5 : 3;
    case SLPB_OP_ADD: case SLPB_OP_SUB: case SLPB_OP_NEG: return 0;
    case SLPB_OP_MUL: return side == 0 ? 2 : 1;
    case SLPB_OP_DIV: return side == 0 ? 2 : 3;
    default: return 3;
  }
}

bool is_sum_op(uint8_t op) {
  return op == SLPB_OP_ADD || op == SLPB_OP_SUB || op == SLPB_OP_NEG;
}

Pattern pattern_from_positions(int32_t rows, int32_t cols,
                               std::vector<std::pair<int32_t, int32_t>>& pos) {
  // pos: (col, row) pairs
  std::sort(pos.begin(), pos.end());
  pos.erase(std::unique(pos.begin(), pos.end()), pos.end());
  Pattern p;
  p.rows = rows;
  p.cols = cols;
  p.colptr.assign(cols + 1, 0);
  p.rowidx.reserve(pos.size());
  for (auto& [c, r] : pos) {
    ++p.colptr[c + 1];
    p.rowidx.push_back(r);
  }
  for (int32_t c = 0; c < cols; ++c) p.colptr[c + 1] += p.colptr[c];
  return p;
}

/// Growing gather description: (entry, source, scale) records in arrival
/// order; finish() groups them by entry with a stable counting sort, so every
/// entry keeps its sources in the order they were added.
struct GatherBuilder {
  struct Record {
    int64_t entry;
    int32_t idx, scale;
  };
  int64_t n_entries;
  std::vector<Record> records;
  explicit GatherBuilder(int64_t n) : n_entries{n} {}
  void add(int64_t entry, int32_t stage, int32_t scale, bool negate = false) {
    records.push_back(
        {entry, negate ? (stage | int32_t(0x80000000)) : stage, scale});
  }
  Gather finish() const {
    Gather g;
    g.ptr.assign(n_entries + 1, 0);
    for (const Record& r : records) ++g.ptr[r.entry + 1];
    for (int64_t e = 0; e < n_entries; ++e) g.ptr[e + 1] += g.ptr[e];
    g.src_idx.resize(records.size());
    g.src_scale.resize(records.size());
    std::vector<int32_t> cursor(g.ptr.begin(), g.ptr.end() - 1);
    for (const Record& r : records) {
      const int32_t at = cursor[r.entry]++;
      g.src_idx[at] = r.idx;
      g.src_scale[at] = r.scale;
    }
    return g;
  }
};

struct Compiler {
  const Tape& tape;
  std::string& error;

  // scratch indexed by node id
  std::vector<int32_t> pos;       // position in the current row list, −1 idle
  std::vector<int32_t> cnt;       // in-row parent count
  std::vector<int32_t> stamp;     // generic visitation stamp
  std::vector<int32_t> local;     // local slot in the current cluster
  std::vector<int32_t> owner;     // first sub-row that touched an interior node
  std::vector<int32_t> outcol;    // output column of a node in the current row
  std::vector<int32_t> term_count;  // split_row: flattened terms under a sum
  // split_row's work lists (one allocation for all rows)
  std::vector<int32_t> sr_order, sr_stack;
  std::vector<std::pair<int32_t, int8_t>> sr_terms, sr_signed_stack;
  std::vector<uint8_t> sr_spine;
  int32_t stamp_counter = 0;

  Compiler(const Tape& t, std::string& err) : tape{t}, error{err} {
    pos.assign(t.n_nodes, -1);
    cnt.assign(t.n_nodes, 0);
    stamp.assign(t.n_nodes, 0);
    local.assign(t.n_nodes, -1);
    outcol.assign(t.n_nodes, -1);
    term_count.assign(t.n_nodes, 0);
  }

  bool is_interior(int32_t node) const { return tape.lhs[node] >= 0; }

  /// Splits `list` (a whole row) into sub-rows at a long root sum.
  /// `make` is called once per sub-row with (seed, nodes).
  template <typename F>
  void split_row(const int32_t* list, int32_t len, F&& make) {
    if (len == 0) return;
    const int32_t root = list[0];
    // in-row parent counts
    for (int32_t i = 0; i < len; ++i) {
      const int32_t nd = list[i];
      if (tape.lhs[nd] >= 0) ++cnt[tape.lhs[nd]];
      if (tape.rhs[nd] >= 0) ++cnt[tape.rhs[nd]];
    }
    // How many terms each expandable sum node would flatten into.
    auto expandable = [&](int32_t nd) {
      return is_sum_op(tape.op[nd]) && (nd == root || cnt[nd] == 1);
    };
    // (node-indexed scratch, 0 = not an expandable sum; reset below)
    std::vector<int32_t>& n_terms = term_count;
    std::vector<int32_t>& order = sr_order;
    order.clear();
    {
      std::vector<int32_t>& st = sr_stack;
      st.assign(1, root);
      while (!st.empty()) {
        const int32_t nd = st.back();
        st.pop_back();
        if (!expandable(nd)) continue;
        order.push_back(nd);
        st.push_back(tape.lhs[nd]);
        if (tape.op[nd] != SLPB_OP_NEG) st.push_back(tape.rhs[nd]);
      }
      for (auto it = order.rbegin(); it != order.rend(); ++it) {
        auto count = [&](int32_t ch) { return std::max(n_terms[ch], 1); };
        int32_t c = count(tape.lhs[*it]);
        if (tape.op[*it] != SLPB_OP_NEG) c += count(tape.rhs[*it]);
        n_terms[*it] = c;
      }
    }
    // The left spine is flattened: ((t0 + t1) + t2) + … summed term by term
    // in that order is the chain itself, so values and the term-major adjoint
    // order of the reference's sweep survive. A right operand stays one term
    // (x + (a + b) is not (x + a) + b) unless it is itself a long sum, which
    // must be cut to keep clusters small; that reassociates it (a few ulp).
    std::vector<std::pair<int32_t, int8_t>>& terms = sr_terms;
    std::vector<std::pair<int32_t, int8_t>>& stack = sr_signed_stack;
    std::vector<uint8_t>& on_spine = sr_spine;
    terms.clear();
    stack.assign(1, {root, int8_t(1)});
    on_spine.assign(1, 1);
    while (!stack.empty()) {
      auto [nd, sg] = stack.back();
      const bool spine = on_spine.back() != 0;
      stack.pop_back();
      on_spine.pop_back();
      const uint8_t op = tape.op[nd];
      if (expandable(nd) && (spine || n_terms[nd] >= kSplitThreshold)) {
        if (op == SLPB_OP_NEG) {
          stack.emplace_back(tape.lhs[nd], int8_t(-sg));
          on_spine.push_back(spine);
        } else {
          // push rhs first so that lhs is expanded first (left-to-right terms)
          stack.emplace_back(tape.rhs[nd],
                             op == SLPB_OP_SUB ? int8_t(-sg) : sg);
          on_spine.push_back(0);
          stack.emplace_back(tape.lhs[nd], sg);
          on_spine.push_back(spine);
        }
      } else {
        terms.emplace_back(nd, sg);
      }
    }
    for (int32_t i = 0; i < len; ++i) cnt[list[i]] = 0;
    for (int32_t nd : order) n_terms[nd] = 0;

    if (static_cast<int>(terms.size()) < kSplitThreshold) {
      make(int8_t(1), std::vector<int32_t>(list, list + len));
      return;
    }
    for (int32_t i = 0; i < len; ++i) pos[list[i]] = i;
    std::vector<int32_t> reach, dfs;
    for (auto [troot, sg] : terms) {
      ++stamp_counter;
      reach.clear();
      dfs.assign(1, troot);
      stamp[troot] = stamp_counter;
      while (!dfs.empty()) {
        int32_t nd = dfs.back();
        dfs.pop_back();
        reach.push_back(nd);
        for (int32_t ch : {tape.lhs[nd], tape.rhs[nd]}) {
          if (ch >= 0 && stamp[ch] != stamp_counter) {
            stamp[ch] = stamp_counter;
            dfs.push_back(ch);
          }
        }
      }
      std::sort(reach.begin(), reach.end(),
                [&](int32_t a, int32_t b) { return pos[a] < pos[b]; });
      make(sg, reach);
    }
    for (int32_t i = 0; i < len; ++i) pos[list[i]] = -1;
  }

  /// Fills sr.flags and drops nodes that are neither active nor value-needed
  /// (e.g. the −y_j term of a Lagrangian-gradient row: it feeds no output and
  /// no partial, and would otherwise chain all time steps into one cluster).
  void analyze_subrow(SubRow& sr) { analyze_subrow(sr, pos); }
  void analyze_subrow(SubRow& sr, std::vector<int32_t>& pos) const {
    const int32_t len = static_cast<int32_t>(sr.nodes.size());
    sr.flags.assign(len, 0);
    if (sr.is_value) {
      std::fill(sr.flags.begin(), sr.flags.end(), kValue);
      return;
    }
    for (int32_t i = 0; i < len; ++i) pos[sr.nodes[i]] = i;
    for (auto& [stage, nd] : sr.outs) sr.flags[pos[nd]] |= kActive;
    for (int32_t i = len - 1; i >= 0; --i) {
      const int32_t nd = sr.nodes[i];
      if (!is_interior(nd)) continue;
      const int32_t l = tape.lhs[nd], r = tape.rhs[nd];
      if ((sr.flags[pos[l]] & kActive) ||
          (r >= 0 && (sr.flags[pos[r]] & kActive))) {
        sr.flags[i] |= kActive;
      }
    }
    for (int32_t i = 0; i < len; ++i) {  // parent→child: one pass closes it
      const int32_t nd = sr.nodes[i];
      if (!is_interior(nd)) continue;
      const int32_t l = tape.lhs[nd], r = tape.rhs[nd];
      if (sr.flags[i] & kValue) {
        sr.flags[pos[l]] |= kValue;
        if (r >= 0) sr.flags[pos[r]] |= kValue;
      }
      if (sr.flags[i] & kActive) {
        for (int side = 0; side < 2; ++side) {
          const int32_t ch = side == 0 ? l : r;
          if (ch < 0 || !(sr.flags[pos[ch]] & kActive)) continue;
          const uint8_t needs = grad_value_needs(tape.op[nd], side);
          if (needs & 1) sr.flags[pos[l]] |= kValue;
          if ((needs & 2) && r >= 0) sr.flags[pos[r]] |= kValue;
          if (needs & 4) sr.flags[i] |= kValue;
        }
      }
    }
    for (int32_t i = 0; i < len; ++i) pos[sr.nodes[i]] = -1;
    int32_t w = 0;
    for (int32_t i = 0; i < len; ++i) {
      if (sr.flags[i] == 0) continue;
      sr.nodes[w] = sr.nodes[i];
      sr.flags[w] = sr.flags[i];
      ++w;
    }
    sr.nodes.resize(w);
    sr.flags.resize(w);
  }

  /// Per-thread scratch indexed by node id (all −1 between uses).
  struct Scratch {
    std::vector<int32_t> pos, local;
  };
  /// What the parallel pass records about one cluster.
  struct ClusterInfo {
    uint64_t h1 = 0, h2 = 0;
    int32_t n_slots = 0;
    bool bad_leaf = false;
    std::vector<int32_t> leaf_index, val_out_stage, adj_out_stage;
    std::vector<double> const_vals;
  };

  /// Runs fn(begin, end, scratch) over [0, n) split across a few host threads
  /// (the compiler's own scratch serves the calling thread).
  template <typename F>
  void parallel_ranges(int32_t n, F&& fn) {
    const unsigned hw = std::thread::hardware_concurrency();
    const int32_t T =
        n < 4096 ? 1 : static_cast<int32_t>(std::min(8u, std::max(1u, hw)));
    Scratch mine;
    mine.pos.swap(pos);
    mine.local.swap(local);
    if (T == 1) {
      fn(0, n, mine);
    } else {
      std::vector<std::thread> threads;
      for (int32_t t = 1; t < T; ++t) {
        threads.emplace_back([&, t] {
          Scratch sc;
          sc.pos.assign(tape.n_nodes, -1);
          sc.local.assign(tape.n_nodes, -1);
          fn(static_cast<int32_t>(int64_t(n) * t / T),
             static_cast<int32_t>(int64_t(n) * (t + 1) / T), sc);
        });
      }
      fn(0, static_cast<int32_t>(int64_t(n) / T), mine);
      for (auto& th : threads) th.join();
    }
    mine.pos.swap(pos);
    mine.local.swap(local);
  }

  /// Local slot numbering (left in sc.local, nodes in cl_nodes), binding data
  /// and signature hashes of one cluster.
  void describe_cluster(const std::vector<SubRow>& subs,
                        const std::vector<int32_t>& mem, Scratch& sc,
                        ClusterInfo& ci, std::vector<int32_t>& cl_nodes) const {
    cl_nodes.clear();
    for (int32_t s : mem) {
      const SubRow& sr = subs[s];
      for (size_t i = 0; i < sr.nodes.size(); ++i) {
        const int32_t nd = sr.nodes[i];
        if ((sr.flags[i] & kValue) && sc.local[nd] < 0) {
          sc.local[nd] = static_cast<int32_t>(cl_nodes.size());
          cl_nodes.push_back(nd);
        }
      }
    }
    ci = ClusterInfo{};
    ci.n_slots = static_cast<int32_t>(cl_nodes.size());
    for (int32_t slot = 0; slot < ci.n_slots; ++slot) {
      const int32_t nd = cl_nodes[slot];
      if (is_interior(nd)) continue;
      if (tape.op[nd] == SLPB_OP_VAR) {
        if (tape.leaf_of_node[nd] < 0) ci.bad_leaf = true;
        ci.leaf_index.push_back(tape.leaf_of_node[nd]);
      } else {
        ci.const_vals.push_back(tape.val[nd]);
      }
    }
    for (int32_t s : mem) {
      const SubRow& sr = subs[s];
      if (sr.is_value) {
        ci.val_out_stage.push_back(sr.value_stage);
      } else {
        for (auto& [stage, nd] : sr.outs) ci.adj_out_stage.push_back(stage);
      }
    }
    uint64_t h1 = 1469598103934665603ull, h2 = 0x9e3779b97f4a7c15ull;
    auto mix = [&](uint32_t w) {
      h1 ^= w;
      h1 *= 1099511628211ull;
      h2 = (h2 ^ (w + 0x9e3779b9u)) * 0xff51afd7ed558ccdull;
      h2 ^= h2 >> 29;
    };
    mix(static_cast<uint32_t>(mem.size()));
    mix(static_cast<uint32_t>(ci.n_slots));
    for (int32_t s : mem) {
      const SubRow& sr = subs[s];
      const int32_t len = static_cast<int32_t>(sr.nodes.size());
      mix(sr.is_value ? 1u : 0u);
      mix(static_cast<uint32_t>(static_cast<int32_t>(sr.seed)));
      mix(static_cast<uint32_t>(len));
      mix(static_cast<uint32_t>(sr.outs.size()));
      for (int32_t i = 0; i < len; ++i) sc.pos[sr.nodes[i]] = i;
      for (int32_t i = 0; i < len; ++i) {
        const int32_t nd = sr.nodes[i];
        const int32_t l = tape.lhs[nd], r = tape.rhs[nd];
        mix(uint32_t(tape.op[nd]) | (uint32_t(sr.flags[i]) << 8));
        mix(static_cast<uint32_t>(sc.local[nd]));
        mix(static_cast<uint32_t>(l >= 0 ? sc.pos[l] : -2));
        mix(static_cast<uint32_t>(r >= 0 ? sc.pos[r] : -2));
      }
      for (auto& [stage, nd] : sr.outs) mix(static_cast<uint32_t>(sc.pos[nd]));
      for (int32_t i = 0; i < len; ++i) sc.pos[sr.nodes[i]] = -1;
    }
    ci.h1 = h1;
    ci.h2 = h2;
  }

  /// Emits the position-independent program of one cluster into `prog`.
  /// local[] must hold the cluster's logical value slots (cl_nodes order).
  ///
  /// Scratch slots are PHYSICAL: values and adjoints share one array and a
  /// slot is recycled once its last reader has run (liveness over the level
  /// schedule: load → forward levels → value outputs → reverse levels → adjoint
  /// outputs). That keeps a cart-pole stage at a few hundred doubles instead of
  /// 1 700, so that 32 stages fit in one CTA's shared memory side by side.
  /// Cost model of the schedule, in units of one multiply-add contribution
  /// (≈140 cycles of a lone warp: two dependent shared-memory loads and the
  /// arithmetic): divisions, square roots and the transcendental functions
  /// run for several hundred cycles (scripts/sweep_debug.py: super-levels of
  /// equal item count took 3.4k … 11.8k cycles).
  static constexpr int32_t kExpensive = 20;
  static bool is_cheap(uint8_t op) {
    switch (op) {
      case SLPB_OP_ADD: case SLPB_OP_SUB: case SLPB_OP_MUL: case SLPB_OP_NEG:
      case SLPB_OP_ABS: case SLPB_OP_SIGN: case SLPB_OP_MAX: case SLPB_OP_MIN:
      case SLPB_OP_IS_NONNEG: case SLPB_OP_IS_POS:
        return true;
      default:
        return false;
    }
  }
  static int32_t forward_cost(uint8_t op) {
    if (flat_costs() || is_cheap(op)) return 1;
    switch (op) {
      case SLPB_OP_DIV:
        return 3;
      case SLPB_OP_POW: case SLPB_OP_ATAN2: case SLPB_OP_HYPOT:
        return 24;
      default:
        return kExpensive;  // sqrt, sin, cos, exp, log, … (FP64, one warp)
    }
  }
  static int32_t contrib_cost(uint8_t op) {
    if (op == kOpLinear || op == kOpLinearNeg || flat_costs()) return 1;
    if (is_cheap(op)) return 1;
    switch (op) {
      case SLPB_OP_DIV:
        return 3;
      case SLPB_OP_SQRT: case SLPB_OP_LOG: case SLPB_OP_LOG10:
        return 5;  // divisions
      case SLPB_OP_POW: case SLPB_OP_ATAN2: case SLPB_OP_HYPOT: case SLPB_OP_TAN:
      case SLPB_OP_TANH:
        return 24;
      default:
        return kExpensive;  // a transcendental function and a multiplication
    }
  }
  static bool flat_costs() {
    static const bool flat = std::getenv("SLPB_SCHED_FLAT_COSTS") != nullptr || legacy_schedule();
    return flat;
  }
  static bool legacy_schedule() {
    static const bool legacy = std::getenv("SLPB_SCHED_LEGACY") != nullptr;
    return legacy;
  }
  /// Set by the caller of emit_program / reported by it.
  int32_t m_window = 1 << 30;
  int64_t m_last_critical_path = 0;
  int32_t m_last_visits = 0;
  template <typename V>
  static int32_t visit_cost(const V& v) {
    int32_t w = 1;
    for (const auto& c : v.contribs) w += contrib_cost(c.op);
    return w;
  }

  bool emit_program(std::vector<SubRow>& subs, const std::vector<int32_t>& mem,
                    const std::vector<int32_t>& cl_nodes,
                    std::vector<int32_t>& level,
                    std::vector<int32_t>& sorted_ids,
                    std::vector<uint32_t>& prog, ProgramSet& ps) {
    (void)ps;
    const int32_t n_slots = static_cast<int32_t>(cl_nodes.size());
    // --- forward levels ------------------------------------------------------
    sorted_ids.assign(cl_nodes.begin(), cl_nodes.end());
    std::sort(sorted_ids.begin(), sorted_ids.end());
    // Forward instruction of every interior slot: operand slots and opcode.
    // The DSL does not share sub-expressions: a dynamics function that writes
    // sin(θ) four times makes four nodes (the cart-pole stage: 32 sin/cos for 8
    // distinct ones, each ≈3 000 cycles of one warp in FP64). A repeated
    // function of the same argument node becomes a COPY of the first one's
    // value (max(v, v) — the same device function on the same argument gives
    // the same bits); SLPB_NO_VALUE_REUSE=1 evaluates every node.
    std::vector<int32_t> f_lhs(n_slots, -1), f_rhs(n_slots, -1);
    std::vector<uint8_t> f_op(n_slots, 0);
    {
      std::map<std::pair<uint8_t, int32_t>, int32_t> first_of;
      for (int32_t nd : sorted_ids) {
        if (!is_interior(nd)) continue;
        const int32_t slot = local[nd];
        f_lhs[slot] = local[tape.lhs[nd]];
        f_rhs[slot] = tape.rhs[nd] >= 0 ? local[tape.rhs[nd]] : -1;
        f_op[slot] = tape.op[nd];
        if (tape.rhs[nd] >= 0 || is_cheap(tape.op[nd]) || tape.op[nd] == SLPB_OP_DIV ||
            !value_reuse_enabled()) {
          continue;
        }
        const auto [it, fresh] =
            first_of.emplace(std::make_pair(tape.op[nd], tape.lhs[nd]), slot);
        if (!fresh) {
          f_lhs[slot] = f_rhs[slot] = it->second;
          f_op[slot] = SLPB_OP_MAX;
        }
      }
    }
    level.assign(n_slots, 0);
    int32_t max_level = 0;
    for (int32_t nd : sorted_ids) {
      if (!is_interior(nd)) continue;
      const int32_t slot = local[nd];
      int32_t lv = level[f_lhs[slot]];
      if (f_rhs[slot] >= 0) lv = std::max(lv, level[f_rhs[slot]]);
      level[slot] = lv + 1;
      max_level = std::max(max_level, lv + 1);
    }
    std::vector<std::vector<int32_t>> fwd_levels(max_level);  // logical slots
    std::vector<int32_t> leaf_slots, const_slots;             // logical slots
    for (int32_t slot = 0; slot < n_slots; ++slot) {
      const int32_t nd = cl_nodes[slot];
      if (is_interior(nd)) {
        fwd_levels[level[slot] - 1].push_back(slot);
      } else if (tape.op[nd] == SLPB_OP_VAR) {
        leaf_slots.push_back(slot);
      } else {
        const_slots.push_back(slot);
      }
    }
    // --- reverse visits ------------------------------------------------------
    struct ContribTmp {
      int32_t parent_visit, l, r;  // l, r: logical value slots or −1
      uint8_t op, side;
    };
    // Partials of +, −, unary − and × are "adjoint × multiplier" with the
    // multiplier +1, −1 or an operand value: a·(±1) is exact, so those
    // contributions are emitted as kOpLinear records (multiplier slot in `l`)
    // that the kernel evaluates without decoding an opcode. ±1 live in two
    // synthetic constant slots (logical ids n_slots, n_slots + 1).
    bool uses_unit_consts = false;
    const int32_t slot_plus1 = n_slots, slot_minus1 = n_slots + 1;
    struct VisitTmp {
      int32_t rlevel, seed;
      std::vector<ContribTmp> contribs;
    };
    std::vector<VisitTmp> visits;
    std::vector<int32_t> adj_out_visit, val_out_slots;
    // (op, argument node) → slot of the sin / cos / sinh / cosh nodes of the
    // cluster
    const bool kNoValueReuse = !value_reuse_enabled();
    std::map<std::pair<uint8_t, int32_t>, int32_t> twins;
    for (int32_t slot = 0; slot < n_slots; ++slot) {
      const int32_t nd = cl_nodes[slot];
      const uint8_t op = tape.op[nd];
      if (op == SLPB_OP_SIN || op == SLPB_OP_COS || op == SLPB_OP_SINH ||
          op == SLPB_OP_COSH) {
        twins.emplace(std::make_pair(op, tape.lhs[nd]), slot);
      }
    }
    for (int32_t s : mem) {
      SubRow& sr = subs[s];
      if (sr.is_value) {
        val_out_slots.push_back(local[sr.nodes[0]]);
        continue;
      }
      const int32_t len = static_cast<int32_t>(sr.nodes.size());
      // one visit per active node, in list order; pos[] = visit index
      for (int32_t i = 0; i < len; ++i) {
        const int32_t nd = sr.nodes[i];
        if (!(sr.flags[i] & kActive)) continue;
        pos[nd] = static_cast<int32_t>(visits.size());
        visits.push_back({0, 0, {}});
      }
      if (pos[sr.nodes[0]] >= 0) visits[pos[sr.nodes[0]]].seed = sr.seed;
      for (int32_t i = 0; i < len; ++i) {
        const int32_t nd = sr.nodes[i];
        if (pos[nd] < 0 || !is_interior(nd)) continue;
        const int32_t l = tape.lhs[nd], r = tape.rhs[nd];
        const int32_t pv = pos[nd];
        for (int side = 0; side < 2; ++side) {
          const int32_t ch = side == 0 ? l : r;
          if (ch < 0 || pos[ch] < 0) continue;
          VisitTmp& cv = visits[pos[ch]];
          const uint8_t needs = grad_value_needs(tape.op[nd], side);
          ContribTmp c{};
          c.parent_visit = pv;
          c.l = (needs & 1) ? local[l] : -1;
          c.r = (needs & 2) ? (r >= 0 ? local[r] : local[l]) : -1;
          c.op = tape.op[nd];
          c.side = static_cast<uint8_t>(side);
          switch (tape.op[nd]) {
            case SLPB_OP_ADD:
              c.op = kOpLinear, c.l = slot_plus1, c.r = -1;
              break;
            case SLPB_OP_SUB:
              c.op = kOpLinear, c.l = side == 0 ? slot_plus1 : slot_minus1;
              c.r = -1;
              break;
            case SLPB_OP_NEG:
              c.op = kOpLinear, c.l = slot_minus1, c.r = -1;
              break;
            case SLPB_OP_MUL:
              c.op = kOpLinear, c.l = side == 0 ? local[r] : local[l];
              c.r = -1;
              break;
            // Partials that the forward sweep of this cluster has already
            // computed as a VALUE: exp(x) itself, cos(x) beside sin(x) and the
            // like. The same device function on the same argument gives the
            // same bits, so a·cos(x) is read from that node's slot instead of
            // running a second FP64 cosine (several hundred cycles of one
            // warp's critical path).
            case SLPB_OP_EXP:
              if (needs & 4) c.op = kOpLinear, c.l = local[nd], c.r = -1;
              break;
            case SLPB_OP_SIN:
            case SLPB_OP_COS:
            case SLPB_OP_SINH:
            case SLPB_OP_COSH: {
              const uint8_t twin_op = tape.op[nd] == SLPB_OP_SIN    ? SLPB_OP_COS
                                      : tape.op[nd] == SLPB_OP_COS  ? SLPB_OP_SIN
                                      : tape.op[nd] == SLPB_OP_SINH ? SLPB_OP_COSH
                                                                    : SLPB_OP_SINH;
              const auto twin = twins.find({twin_op, l});
              if (!kNoValueReuse && twin != twins.end()) {
                c.op = tape.op[nd] == SLPB_OP_COS ? kOpLinearNeg : kOpLinear;
                c.l = twin->second, c.r = -1;
              }
              break;
            }
            default:
              break;
          }
          if (c.op == kOpLinear && c.l >= n_slots) uses_unit_consts = true;
          cv.contribs.push_back(c);
          cv.rlevel = std::max(cv.rlevel, visits[pv].rlevel + 1);
        }
      }
      for (auto& [stage, nd] : sr.outs) adj_out_visit.push_back(pos[nd]);
      for (int32_t i = 0; i < len; ++i) pos[sr.nodes[i]] = -1;
    }
    const int32_t n_visits = static_cast<int32_t>(visits.size());

    // --- worker chains ---------------------------------------------------------
    // A block of the instruction stream is not one dependency level but a
    // SUPER-LEVEL: every worker (a row of LC threads, lane = cluster) runs a
    // private list of items in order, and the thread block synchronises only
    // between super-levels. A thread reads back what it wrote itself in program
    // order, so an item whose newest operands were all produced by ONE worker
    // in the current super-level joins that worker's list and needs no barrier;
    // only an item that joins results of different workers waits for the next
    // one. A direct-transcription stage (four RK4 evaluations of the dynamics,
    // ~145 dependency levels) collapses to a few dozen barriers this way.
    int32_t widest = 1;
    for (const auto& lv : fwd_levels) widest = std::max<int32_t>(widest, lv.size());
    {
      std::vector<int32_t> cnt;
      for (const auto& v : visits) {
        if (v.rlevel >= static_cast<int32_t>(cnt.size())) cnt.resize(v.rlevel + 1, 0);
        widest = std::max(widest, ++cnt[v.rlevel]);
      }
    }
    int32_t n_workers = 1;
    while (n_workers < 16 && n_workers < widest) n_workers <<= 1;
    // weight a worker takes per super-level (development switch SLPB_CHAIN_CAP)
    static const int32_t kChainCap = [] {
      const char* e = std::getenv("SLPB_CHAIN_CAP");
      return e ? std::max(1, std::atoi(e)) : 24;
    }();
    const bool kLegacySchedule = legacy_schedule();
    struct Loads {
      int32_t W;
      std::vector<std::vector<int32_t>> load;  // [super-level][worker]
      int32_t& at(int32_t sl, int32_t w) {
        if (sl >= static_cast<int32_t>(load.size())) {
          load.resize(sl + 1, std::vector<int32_t>(W, 0));
        }
        return load[sl][w];
      }
      int32_t lightest(int32_t sl) {
        at(sl, 0);
        int32_t best = 0;
        for (int32_t w = 1; w < W; ++w) {
          if (load[sl][w] < load[sl][best]) best = w;
        }
        return best;
      }
    };
    // dep = (super-level, worker) of every producer this item reads; returns
    // where the item goes
    auto place = [&](Loads& L, const std::pair<int32_t, int32_t>* deps, int n_deps,
                     int32_t first_sl, int32_t weight) {
      int32_t s_max = -1;
      for (int k = 0; k < n_deps; ++k) s_max = std::max(s_max, deps[k].first);
      int32_t sl, w;
      if (s_max < first_sl) {
        sl = first_sl;  // reads only what was there before the sweep began
        w = L.lightest(sl);
      } else {
        int32_t same = -2;  // −2: none yet, −1: several workers
        for (int k = 0; k < n_deps; ++k) {
          if (deps[k].first != s_max) continue;
          same = same == -2 ? deps[k].second
                            : (same == deps[k].second ? same : -1);
        }
        if (same >= 0 && L.at(s_max, same) + weight <= kChainCap) {
          sl = s_max;  // continues that worker's chain: no barrier
          w = same;
        } else {
          // behind a barrier anyway: any worker may take it. The lightest one
          // does (stacking it on its producer's worker left most workers idle
          // in the deep super-levels: critical path 467 → 341 weight units on
          // the cart-pole stage together with the priority order below).
          // Development switch SLPB_SCHED_LEGACY=1: the previous placement.
          sl = s_max + 1;
          if (!kLegacySchedule) {
            w = L.lightest(sl);
          } else {
            w = same >= 0 ? same : -1;
            if (w < 0 || L.at(sl, w) + weight > kChainCap) {
              w = -1;
              for (int k = 0; k < n_deps; ++k) {
                if (deps[k].first != s_max) continue;
                if (w < 0 || L.at(sl, deps[k].second) < L.at(sl, w)) w = deps[k].second;
              }
              if (L.at(sl, w) + weight > kChainCap) w = L.lightest(sl);
            }
          }
        }
      }
      L.at(sl, w) += weight;
      return std::pair<int32_t, int32_t>{sl, w};
    };
    std::vector<int32_t> fwd_worker(n_slots, -1);
    {
      Loads L{n_workers, {}};
      std::vector<int32_t> sl_of(n_slots, 0);  // 0: leaf / constant
      for (int32_t nd : sorted_ids) {
        if (!is_interior(nd)) continue;
        const int32_t slot = local[nd];
        std::pair<int32_t, int32_t> deps[2];
        int n_deps = 0;
        const int32_t a = f_lhs[slot];
        if (sl_of[a] > 0) deps[n_deps++] = {sl_of[a], fwd_worker[a]};
        if (f_rhs[slot] >= 0 && f_rhs[slot] != a) {
          const int32_t b = f_rhs[slot];
          if (sl_of[b] > 0) deps[n_deps++] = {sl_of[b], fwd_worker[b]};
        }
        const auto [sl, w] = place(L, deps, n_deps, 1, forward_cost(f_op[slot]));
        sl_of[slot] = sl;
        fwd_worker[slot] = w;
      }
      max_level = 0;
      for (int32_t slot = 0; slot < n_slots; ++slot) {
        level[slot] = sl_of[slot];
        max_level = std::max(max_level, sl_of[slot]);
      }
      fwd_levels.assign(max_level, {});
      for (int32_t nd : sorted_ids) {  // topological order inside every list
        if (is_interior(nd)) fwd_levels[level[local[nd]] - 1].push_back(local[nd]);
      }
      for (auto& lv : fwd_levels) {
        std::stable_sort(lv.begin(), lv.end(), [&](int32_t x, int32_t y) {
          return fwd_worker[x] < fwd_worker[y];
        });
      }
    }
    std::vector<int32_t> rev_worker(n_visits, -1);
    {
      Loads L{n_workers, {}};
      std::vector<std::pair<int32_t, int32_t>> deps;
      auto place_visit = [&](int32_t i) {
        deps.clear();
        for (const ContribTmp& c : visits[i].contribs) {
          deps.push_back({visits[c.parent_visit].rlevel, rev_worker[c.parent_visit]});
        }
        const int32_t weight = visit_cost(visits[i]);
        const auto [sl, w] =
            place(L, deps.data(), static_cast<int>(deps.size()), 0, weight);
        visits[i].rlevel = sl;
        rev_worker[i] = w;
      };
      if (kLegacySchedule) {
        for (int32_t i = 0; i < n_visits; ++i) place_visit(i);  // parents first
      } else {
        // list scheduling: among the visits whose parents are placed, the one
        // with the longest remaining dependency chain goes first
        std::vector<int64_t> bottom(n_visits, 0);
        std::vector<int32_t> child_ptr(n_visits + 1, 0), pending(n_visits, 0);
        for (int32_t i = 0; i < n_visits; ++i) {
          for (const ContribTmp& c : visits[i].contribs) ++child_ptr[c.parent_visit + 1];
          pending[i] = static_cast<int32_t>(visits[i].contribs.size());
        }
        for (int32_t i = 0; i < n_visits; ++i) child_ptr[i + 1] += child_ptr[i];
        std::vector<int32_t> child(child_ptr[n_visits]), fill(child_ptr.begin(), child_ptr.end() - 1);
        for (int32_t i = 0; i < n_visits; ++i) {
          for (const ContribTmp& c : visits[i].contribs) child[fill[c.parent_visit]++] = i;
        }
        for (int32_t i = n_visits - 1; i >= 0; --i) {
          int64_t below = 0;
          for (int32_t k = child_ptr[i]; k < child_ptr[i + 1]; ++k) below = std::max(below, bottom[child[k]]);
          bottom[i] = below + visit_cost(visits[i]);
        }
        // … among those at most kWindow visits ahead of the oldest unplaced
        // one: running far ahead along the critical chains keeps the operands
        // of everything left behind alive (60 more scratch slots on the
        // cart-pole stage, which then no longer fits twice on an SM)
        const int32_t kWindow = m_window;
        using Key = std::pair<int64_t, int32_t>;  // (−bottom, index): smallest first
        std::priority_queue<Key, std::vector<Key>, std::greater<Key>> ready;
        std::priority_queue<int32_t, std::vector<int32_t>, std::greater<int32_t>> waiting;
        std::vector<uint8_t> placed(n_visits, 0);
        int32_t lo = 0;  // oldest unplaced visit
        auto offer = [&](int32_t i) {
          if (i < lo + kWindow) {
            ready.push({-bottom[i], i});
          } else {
            waiting.push(i);
          }
        };
        for (int32_t i = 0; i < n_visits; ++i) {
          if (pending[i] == 0) offer(i);
        }
        int32_t n_placed = 0;
        while (n_placed < n_visits) {
          // (the oldest unplaced visit is always ready or behind a ready one,
          // so `ready` cannot run dry before everything is placed)
          const int32_t i = ready.top().second;
          ready.pop();
          place_visit(i);
          placed[i] = 1;
          ++n_placed;
          for (int32_t k = child_ptr[i]; k < child_ptr[i + 1]; ++k) {
            if (--pending[child[k]] == 0) offer(child[k]);
          }
          if (i == lo) {
            while (lo < n_visits && placed[lo]) ++lo;
            while (!waiting.empty() && waiting.top() < lo + kWindow) {
              ready.push({-bottom[waiting.top()], waiting.top()});
              waiting.pop();
            }
          }
        }
      }
    }
    int32_t max_rlevel = -1;
    for (auto& v : visits) max_rlevel = std::max(max_rlevel, v.rlevel);
    std::vector<std::vector<int32_t>> rev_levels(max_rlevel + 1);
    for (int32_t i = 0; i < n_visits; ++i) {
      rev_levels[visits[i].rlevel].push_back(i);
    }
    for (auto& lv : rev_levels) {
      std::stable_sort(lv.begin(), lv.end(), [&](int32_t x, int32_t y) {
        return rev_worker[x] < rev_worker[y];
      });
    }
    {
      // critical path of the schedule: Σ over super-levels of the heaviest list
      const bool report = std::getenv("SLPB_COMPILE_TIMING") != nullptr;
      int64_t crit = 0, total = 0;
      for (const auto& lv : rev_levels) {
        std::vector<int64_t> load(n_workers, 0);
        for (int32_t i : lv) load[rev_worker[i]] += visit_cost(visits[i]);
        crit += *std::max_element(load.begin(), load.end());
        for (int64_t l : load) total += l;
        if (report && std::getenv("SLPB_SCHEDULE_DUMP") && n_workers == 16) {
          std::fprintf(stderr, "   level %2d:", int(&lv - &rev_levels[0]));
          for (int64_t l : load) std::fprintf(stderr, " %3lld", (long long)l);
          std::fprintf(stderr, "\n");
        }
      }
      int64_t fcrit = 0, ftotal = 0;
      for (const auto& lv : fwd_levels) {
        std::vector<int64_t> load(n_workers, 0);
        for (int32_t slot : lv) load[fwd_worker[slot]] += forward_cost(f_op[slot]);
        if (report && std::getenv("SLPB_SCHEDULE_DUMP") && n_workers == 16) {
          std::fprintf(stderr, "   forward %2d:", int(&lv - &fwd_levels[0]));
          for (int64_t l : load) std::fprintf(stderr, " %3lld", (long long)l);
          std::fprintf(stderr, "\n");
        }
        fcrit += *std::max_element(load.begin(), load.end());
        ftotal += static_cast<int64_t>(lv.size());
      }
      if (report && n_workers == 16) {
        std::map<int, int> fh, ch;
        for (const auto& lv : fwd_levels) {
          for (int32_t slot : lv) ++fh[f_op[slot]];
        }
        for (const auto& v : visits) {
          for (const auto& c : v.contribs) ++ch[c.op * 2 + c.side];
        }
        std::fprintf(stderr, "[slpb compile] forward ops:");
        for (auto& [op, n] : fh) std::fprintf(stderr, " %d×%d", op, n);
        std::fprintf(stderr, " | contributions (op.side):");
        for (auto& [k, n] : ch) std::fprintf(stderr, " %d.%d×%d", k / 2, k % 2, n);
        std::fprintf(stderr, "\n");
      }
      m_last_critical_path = fcrit + crit;
      m_last_visits = n_visits;
      if (report) {
        // longest weighted dependency chain of the reverse sweep: no schedule
        // can be shorter
        int64_t dag = 0;
        std::vector<int64_t> cp(n_visits, 0);
        for (int32_t i = 0; i < n_visits; ++i) {
          int64_t in = 0;
          for (const ContribTmp& c : visits[i].contribs) in = std::max(in, cp[c.parent_visit]);
          cp[i] = in + visit_cost(visits[i]);
          dag = std::max(dag, cp[i]);
        }
        std::fprintf(stderr,
                     "[slpb compile] schedule (window %d): %d workers, forward %zu super-levels, "
                     "weight %lld, critical path %lld; reverse %zu super-levels, weight %lld, "
                     "critical path %lld (perfect balance %lld, longest dependency chain %lld)\n",
                     m_window, n_workers, fwd_levels.size(), (long long)ftotal, (long long)fcrit,
                     rev_levels.size(), (long long)total, (long long)crit,
                     (long long)(total / n_workers), (long long)dag);
      }
    }

    // --- physical slot allocation (liveness over the level schedule) ---------
    // time 0: leaf/constant load; 1..Lf: forward levels; Lf+1: value outputs;
    // Lf+2+r: reverse level r; t_end: adjoint outputs.
    const int32_t Lf = max_level;
    const int32_t t_valout = Lf + 1;
    const int32_t t_end = Lf + 2 + (max_rlevel + 1);
    auto t_rev = [&](int32_t rl) { return Lf + 2 + rl; };
    const int32_t n_synth = uses_unit_consts ? 2 : 0;
    const int32_t n_val = n_slots + n_synth;  // value-like items
    level.resize(n_val, 0);  // synthetic constants are loaded at time 0
    // item id: value slot s → s ; visit i → n_val + i. For every item: when it
    // is read last, and — if all its readers at that time sit on ONE worker —
    // by which worker and at which position of that worker's list.
    const int32_t n_items = n_val + n_visits;
    std::vector<int32_t> last_t(n_items, 0), last_w(n_items, -1), last_pos(n_items, 0);
    std::vector<int32_t> fwd_pos(n_slots, 0), rev_pos(n_visits, 0);
    for (auto& lv : fwd_levels) {
      std::vector<int32_t> next(n_workers, 0);
      for (int32_t slot : lv) fwd_pos[slot] = next[fwd_worker[slot]]++;
    }
    for (auto& lv : rev_levels) {
      std::vector<int32_t> next(n_workers, 0);
      for (int32_t vi : lv) rev_pos[vi] = next[rev_worker[vi]]++;
    }
    auto note = [&](int32_t item, int32_t t, int32_t w, int32_t pos) {
      if (t > last_t[item]) {
        last_t[item] = t;
        last_w[item] = w;
        last_pos[item] = pos;
      } else if (t == last_t[item]) {
        if (last_w[item] != w) {
          last_w[item] = -1;
        } else {
          last_pos[item] = std::max(last_pos[item], pos);
        }
      }
    };
    for (int32_t slot = 0; slot < n_slots; ++slot) {  // an item "reads" itself
      if (is_interior(cl_nodes[slot])) {
        note(slot, level[slot], fwd_worker[slot], fwd_pos[slot]);
      }
    }
    for (int32_t i = 0; i < n_visits; ++i) {
      note(n_val + i, t_rev(visits[i].rlevel), rev_worker[i], rev_pos[i]);
    }
    for (int32_t slot = 0; slot < n_slots; ++slot) {
      const int32_t nd = cl_nodes[slot];
      if (!is_interior(nd)) continue;
      note(f_lhs[slot], level[slot], fwd_worker[slot], fwd_pos[slot]);
      if (f_rhs[slot] >= 0) {
        note(f_rhs[slot], level[slot], fwd_worker[slot], fwd_pos[slot]);
      }
    }
    for (int32_t slot : val_out_slots) note(slot, t_valout, -1, 0);
    for (int32_t i = 0; i < n_visits; ++i) {
      const int32_t t = t_rev(visits[i].rlevel);
      for (const ContribTmp& c : visits[i].contribs) {
        note(n_val + c.parent_visit, t, rev_worker[i], rev_pos[i]);
        if (c.l >= 0) note(c.l, t, rev_worker[i], rev_pos[i]);
        if (c.r >= 0) note(c.r, t, rev_worker[i], rev_pos[i]);
      }
    }
    for (int32_t vi : adj_out_visit) note(n_val + vi, t_end, -1, 0);

    // Allocation in execution order. A slot whose last reader runs at time t
    // is reusable from t + 1 on (other workers may still be reading it during
    // t) — unless everything that reads it at time t belongs to one worker:
    // then that worker may take it again right after its last read (program
    // order protects it), which is what keeps the temporaries of a long
    // private chain from piling up.
    std::vector<int32_t> phys(n_items, -1);
    std::vector<int32_t> free_list;  // min-heap of slots any worker may take
    int32_t n_scratch = 0;
    auto heap_cmp = std::greater<int32_t>{};
    auto heap_push = [&](std::vector<int32_t>& h, int32_t v) {
      h.push_back(v);
      std::push_heap(h.begin(), h.end(), heap_cmp);
    };
    auto heap_pop = [&](std::vector<int32_t>& h) {
      std::pop_heap(h.begin(), h.end(), heap_cmp);
      const int32_t v = h.back();
      h.pop_back();
      return v;
    };
    std::vector<std::vector<int32_t>> shared_frees(t_end + 1);
    for (int32_t item = 0; item < n_items; ++item) {
      if (last_w[item] < 0) shared_frees[last_t[item]].push_back(item);
    }
    // time 0: leaves, constants, synthetic constants
    for (int32_t slot = 0; slot < n_val; ++slot) {
      if (slot >= n_slots || !is_interior(cl_nodes[slot])) phys[slot] = n_scratch++;
    }
    std::vector<int32_t> pending, local_free, operands;
    auto run_lists = [&](int32_t t, const std::vector<int32_t>& list, bool reverse) {
      // `list` is grouped by worker, each group in execution order
      size_t k = 0;
      while (k < list.size()) {
        const int32_t w = reverse ? rev_worker[list[k]] : fwd_worker[list[k]];
        local_free.clear();
        int32_t pos = 0;
        for (; k < list.size() &&
               (reverse ? rev_worker[list[k]] : fwd_worker[list[k]]) == w;
             ++k, ++pos) {
          const int32_t x = list[k];
          const int32_t item = reverse ? n_val + x : x;
          phys[item] = !local_free.empty() ? heap_pop(local_free)
                       : !free_list.empty() ? heap_pop(free_list)
                                            : n_scratch++;
          operands.clear();
          operands.push_back(item);
          if (reverse) {
            for (const ContribTmp& c : visits[x].contribs) {
              operands.push_back(n_val + c.parent_visit);
              if (c.l >= 0) operands.push_back(c.l);
              if (c.r >= 0) operands.push_back(c.r);
            }
          } else {
            operands.push_back(f_lhs[x]);
            if (f_rhs[x] >= 0) operands.push_back(f_rhs[x]);
          }
          std::sort(operands.begin(), operands.end());
          operands.erase(std::unique(operands.begin(), operands.end()), operands.end());
          for (int32_t o : operands) {
            if (last_t[o] == t && last_w[o] == w && last_pos[o] == pos) {
              heap_push(local_free, phys[o]);
            }
          }
        }
        pending.insert(pending.end(), local_free.begin(), local_free.end());
      }
    };
    for (int32_t t = 0; t <= t_end; ++t) {
      pending.clear();
      if (t >= 1 && t <= Lf) {
        run_lists(t, fwd_levels[t - 1], false);
      } else if (t >= Lf + 2 && t < t_end) {
        run_lists(t, rev_levels[t - Lf - 2], true);
      }
      for (int32_t item : shared_frees[t]) pending.push_back(phys[item]);
      for (int32_t v : pending) heap_push(free_list, v);
    }
    if (n_scratch > 65535) {
      error = "an expression cluster needs more than 65535 scratch slots; the "
              "block-cooperative fallback for such graphs is not implemented";
      return false;
    }
    if (n_scratch == 0) n_scratch = 1;
    auto vp = [&](int32_t slot) {
      return static_cast<uint16_t>(slot >= 0 ? phys[slot] : 0);
    };
    auto ap = [&](int32_t visit) {
      return static_cast<uint16_t>(phys[n_val + visit]);
    };
    for (int32_t k = 0; k < n_synth; ++k) const_slots.push_back(n_slots + k);

    // --- emit ----------------------------------------------------------------
    prog.assign(kProgHeaderWords, 0);
    auto align4 = [&] {
      while (prog.size() & 3) prog.push_back(0);
    };
    auto push_u16s = [&](const std::vector<uint16_t>& v) {
      const size_t off = prog.size();
      prog.resize(off + (v.size() + 1) / 2, 0);
      if (!v.empty()) {
        std::memcpy(prog.data() + off, v.data(), v.size() * sizeof(uint16_t));
      }
      return static_cast<uint32_t>(off);
    };
    auto phys_list = [&](const std::vector<int32_t>& logical, bool adjoint) {
      std::vector<uint16_t> out;
      out.reserve(logical.size());
      for (int32_t x : logical) out.push_back(adjoint ? ap(x) : vp(x));
      return out;
    };
    auto push_record = [&](const void* rec) {
      uint32_t w[2];
      std::memcpy(w, rec, 8);
      prog.push_back(w[0]);
      prog.push_back(w[1]);
    };
    const bool has_valout = !val_out_slots.empty();
    const int32_t n_blocks = static_cast<int32_t>(fwd_levels.size()) +
                             (has_valout ? 1 : 0) +
                             static_cast<int32_t>(rev_levels.size());
    prog[0] = static_cast<uint32_t>(n_scratch);
    prog[1] = static_cast<uint32_t>(n_slots + n_visits);  // logical, for stats
    prog[2] = static_cast<uint32_t>(leaf_slots.size());
    prog[3] = static_cast<uint32_t>(const_slots.size());
    prog[4] = static_cast<uint32_t>(n_blocks);
    prog[6] = static_cast<uint32_t>(val_out_slots.size());
    prog[7] = static_cast<uint32_t>(adj_out_visit.size());
    // tables (copied to shared memory together with the header)
    prog[8] = push_u16s(phys_list(leaf_slots, false));
    prog[9] = push_u16s(phys_list(const_slots, false));
    prog[15] = push_u16s(phys_list(val_out_slots, false));
    prog[16] = push_u16s(phys_list(adj_out_visit, true));
    prog[10] = static_cast<uint32_t>(prog.size());
    const size_t table_at = prog.size();
    prog.resize(prog.size() + n_blocks + 1, 0);
    align4();
    prog[5] = static_cast<uint32_t>(prog.size());  // prologue words
    // level blocks, each 16-byte aligned: {kind, n_items, n_contrib, 0} + payload
    int32_t max_width = 1, blk = 0;
    uint32_t max_block_words = 4, n_instr = 0, n_contrib_total = 0;
    // {kind, n_items, n_contrib, W} then, for W > 0, W + 1 item offsets (one
    // list per worker) padded to an even number of words
    auto begin_block = [&](uint32_t kind, uint32_t items, uint32_t contribs,
                           const std::vector<int32_t>* list = nullptr,
                           const std::vector<int32_t>* worker_of = nullptr) {
      align4();
      prog[table_at + blk] = static_cast<uint32_t>(prog.size());
      prog.push_back(kind);
      prog.push_back(items);
      prog.push_back(contribs);
      prog.push_back(list ? static_cast<uint32_t>(n_workers) : 0u);
      if (list) {
        std::vector<uint32_t> off(n_workers + 1, 0);
        for (int32_t x : *list) ++off[(*worker_of)[x] + 1];
        for (int32_t w = 0; w < n_workers; ++w) off[w + 1] += off[w];
        prog.insert(prog.end(), off.begin(), off.end());
        if ((n_workers + 1) & 1) prog.push_back(0);
      }
    };
    auto end_block = [&] {
      align4();
      max_block_words = std::max<uint32_t>(
          max_block_words,
          static_cast<uint32_t>(prog.size()) - prog[table_at + blk]);
      ++blk;
    };
    for (auto& lv : fwd_levels) {
      begin_block(kBlockForward, static_cast<uint32_t>(lv.size()), 0, &lv,
                  &fwd_worker);
      max_width = std::max<int32_t>(max_width, lv.size());
      for (int32_t slot : lv) {
        FwdInstr in{};
        in.dst = vp(slot);
        in.a = vp(f_lhs[slot]);
        in.b = f_rhs[slot] >= 0 ? vp(f_rhs[slot]) : in.a;
        in.op = f_op[slot];
        push_record(&in);
        ++n_instr;
      }
      end_block();
    }
    if (has_valout) {
      begin_block(kBlockValueOut, 0, 0);
      end_block();
    }
    for (auto& lv : rev_levels) {
      uint32_t nc = 0;
      for (int32_t vi : lv) nc += static_cast<uint32_t>(visits[vi].contribs.size());
      begin_block(kBlockReverse, static_cast<uint32_t>(lv.size()), nc, &lv,
                  &rev_worker);
      max_width = std::max<int32_t>(max_width, lv.size());
      uint32_t crun = 0;
      for (int32_t vi : lv) {
        const VisitTmp& v = visits[vi];
        Visit rec{};
        rec.adj = ap(vi);
        if (v.contribs.empty()) {
          // a root (seeded) visit; an unseeded node without parents cannot
          // occur because every non-root node of a list has a parent in it
          rec.n_contrib = 0;
          rec.seed = static_cast<int8_t>(v.seed);
          rec.contrib_begin = 0;
        } else {
          if (v.contribs.size() > 254) {
            error = "a node has more than 254 parents inside one row";
            return false;
          }
          rec.n_contrib = static_cast<uint8_t>(v.contribs.size());
          rec.seed = 0;
          rec.contrib_begin = crun;  // relative to the block's contributions
          crun += static_cast<uint32_t>(v.contribs.size());
        }
        push_record(&rec);
      }
      for (int32_t vi : lv) {
        for (const ContribTmp& c : visits[vi].contribs) {
          Contrib rec{};
          rec.parent_adj = ap(c.parent_visit);
          // An operand the partial does not use still gets loaded by the
          // interpreter (unconditional loads keep the inner loop branch-free).
          // Point it at the parent's adjoint slot: live and not written during
          // this level, unlike physical slot 0, which a visit of the same
          // level may own (compute-sanitizer racecheck flagged that dead read).
          rec.l = c.l >= 0 ? vp(c.l) : rec.parent_adj;
          rec.r = c.r >= 0 ? vp(c.r) : rec.parent_adj;
          rec.op = c.op;
          rec.side = c.side;
          push_record(&rec);
        }
      }
      n_contrib_total += nc;
      end_block();
    }
    align4();
    prog[table_at + n_blocks] = static_cast<uint32_t>(prog.size());
    prog[11] = max_block_words;
    prog[12] = static_cast<uint32_t>(fwd_levels.size());
    prog[13] = static_cast<uint32_t>(rev_levels.size());
    prog[14] = n_instr;
    prog[17] = static_cast<uint32_t>(max_width);
    prog[21] = static_cast<uint32_t>(n_workers);
    {
      const uint32_t stream_words =
          prog[table_at + n_blocks] - prog[table_at];
      const uint32_t with_stream =
          static_cast<uint32_t>(n_scratch) * 32u * 8u + prog[5] * 4u +
          stream_words * 4u + 64u;
      const bool resident =
          stream_words * 4u <= kAdResidentBytes && with_stream <= kAdHalfSmBytes;
      prog[22] = resident ? stream_words : kAdStages * max_block_words;
      prog[23] = resident ? 1u : 0u;
      if (std::getenv("SLPB_COMPILE_TIMING") && n_workers == 16) {
        std::fprintf(stderr,
                     "[slpb compile] %d scratch slots (%u B at 32 lanes), tables %u B, "
                     "stream %u B in %d blocks (largest %u B), %s\n",
                     n_scratch, static_cast<uint32_t>(n_scratch) * 256u, prog[5] * 4u,
                     stream_words * 4u, n_blocks, max_block_words * 4u,
                     resident ? "resident" : "ring");
      }
    }
    prog[18] = n_contrib_total;
    prog[19] = static_cast<uint32_t>(n_visits);
    prog[20] = static_cast<uint32_t>(n_synth);  // trailing ±1 constants
    return true;
  }

  /// Builds clusters + programs for `subs`.
  bool build_programs(std::vector<SubRow>& subs, ProgramSet& ps) {
    const int32_t ns = static_cast<int32_t>(subs.size());
    StageTimer timer;
    // --- clusters: union-find over sub-rows that share an interior node ----
    std::vector<int32_t> uf(ns);
    std::iota(uf.begin(), uf.end(), 0);
    auto find = [&](int32_t a) {
      while (uf[a] != a) a = uf[a] = uf[uf[a]];
      return a;
    };
    owner.assign(tape.n_nodes, -1);
    parallel_ranges(ns, [&](int32_t b, int32_t e, Scratch& sc) {
      for (int32_t s = b; s < e; ++s) analyze_subrow(subs[s], sc.pos);
    });
    timer.lap("analyze_subrow");
    // Fusing is an optimisation (shared nodes are evaluated once); clusters
    // that stay apart each evaluate their own copy, as the reference does per
    // row. Without a bound, rows that overlap pairwise — the Lagrangian
    // gradient rows of a collocation OCP, where every state enters two
    // consecutive defects nonlinearly — chain ALL time steps into one cluster
    // that no thread block can hold.
    std::vector<int64_t> weight(ns);
    for (int32_t s = 0; s < ns; ++s) {
      weight[s] = static_cast<int64_t>(subs[s].nodes.size());
    }
    for (int32_t s = 0; s < ns; ++s) {
      const SubRow& sr = subs[s];
      for (size_t i = 0; i < sr.nodes.size(); ++i) {
        const int32_t nd = sr.nodes[i];
        if (!is_interior(nd) || !(sr.flags[i] & kValue)) continue;
        if (owner[nd] < 0) {
          owner[nd] = s;
        } else {
          const int32_t a = find(owner[nd]), b = find(s);
          if (a == b) continue;
          if (weight[a] + weight[b] > kMaxClusterVisits) {
            owner[nd] = s;  // later rows sharing nd may still join this one
            continue;
          }
          const int32_t root = std::min(a, b), child = std::max(a, b);
          uf[child] = root;
          weight[root] += weight[child];
        }
      }
    }
    std::vector<int32_t> cluster_of(ns);
    std::vector<std::vector<int32_t>> members;
    {
      std::vector<int32_t> cid(ns, -1);
      for (int32_t s = 0; s < ns; ++s) {
        int32_t r = find(s);
        if (cid[r] < 0) {
          cid[r] = static_cast<int32_t>(members.size());
          members.emplace_back();
        }
        cluster_of[s] = cid[r];
        members[cid[r]].push_back(s);
      }
    }

    timer.lap("union-find + members");
    if (timer.on) {
      size_t biggest = 0;
      for (const auto& m : members) {
        size_t total = 0;
        for (int32_t sub : m) total += subs[sub].nodes.size();
        biggest = std::max(biggest, total);
      }
      std::fprintf(stderr,
                   "[slpb compile] %d sub-rows in %zu clusters, largest %zu "
                   "node visits\n",
                   ns, members.size(), biggest);
    }
    // --- per cluster, in parallel: local slot numbering, binding data and a
    // 128-bit hash of the structural signature (everything the program depends
    // on, in cluster-local terms: slot numbers, positions inside each sub-row).
    // Clusters with equal hashes share one program — every time step after the
    // first finds its program without it being rebuilt. (Two independent 64-bit
    // hashes; a collision between different signatures is a 2⁻¹²⁸ event.)
    const int32_t n_clusters = static_cast<int32_t>(members.size());
    std::vector<ClusterInfo> infos(n_clusters);
    parallel_ranges(n_clusters, [&](int32_t b, int32_t e, Scratch& sc) {
      std::vector<int32_t> cl_nodes;
      for (int32_t c = b; c < e; ++c) {
        describe_cluster(subs, members[c], sc, infos[c], cl_nodes);
        for (int32_t nd : cl_nodes) sc.local[nd] = -1;
      }
    });
    for (const ClusterInfo& ci : infos) {
      if (ci.n_slots > 65535) {
        error = "an expression cluster has more than 65535 nodes; the block-"
                "cooperative fallback for such graphs is not implemented";
        return false;
      }
      if (ci.bad_leaf) {
        error = "the tape contains a decision-variable node that is not a "
                "decision variable, y multiplier or z multiplier";
        return false;
      }
    }
    timer.lap("signatures (parallel)");

    // --- sequential merge: programs in first-seen order, bindings appended ----
    std::map<std::pair<uint64_t, uint64_t>, int32_t> by_hash;
    std::vector<uint32_t> prog;       // program being built
    std::vector<int32_t> cl_nodes;    // nodes of the cluster, first-seen order
    std::vector<int32_t> level;       // per local slot
    std::vector<int32_t> sorted_ids;
    Scratch main_sc;
    main_sc.pos.assign(tape.n_nodes, -1);
    main_sc.local.assign(tape.n_nodes, -1);
    std::fill(local.begin(), local.end(), -1);
    for (int32_t c = 0; c < n_clusters; ++c) {
      const auto& mem = members[c];
      ClusterInfo& ci = infos[c];
      const auto key = std::make_pair(ci.h1, ci.h2);
      auto it = by_hash.find(key);
      int32_t pid;
      if (it != by_hash.end()) {
        pid = it->second;
      } else {
        // a new program: rebuild the cluster's local numbering in the
        // compiler's own scratch (emit_program reads `local`) and emit it
        ClusterInfo again;
        describe_cluster(subs, mem, main_sc, again, cl_nodes);
        for (size_t k = 0; k < cl_nodes.size(); ++k) {
          local[cl_nodes[k]] = static_cast<int32_t>(k);
        }
        // The reverse sweep is list-scheduled by remaining chain length inside
        // a window of the visit order (emit_program): a wide window shortens
        // the critical path, a narrow one keeps fewer operands alive. Take the
        // shortest schedule among those whose task (32 lanes) still fits twice
        // on an SM; SLPB_SCHED_WINDOW pins the window (development).
        bool ok = true;
        {
          static const int32_t kPinned = [] {
            const char* e = std::getenv("SLPB_SCHED_WINDOW");
            return e ? std::max(1, std::atoi(e)) : 0;
          }();
          const int32_t windows[] = {1 << 30, 512, 256, 128, 64, 32};
          std::vector<uint32_t> best;
          int64_t best_cost = -1;
          bool best_fits = false;
          for (int32_t window : windows) {
            // (a window that covers the whole sweep is the unbounded one)
            if (best_cost >= 0 && window >= m_last_visits) continue;
            m_window = kPinned > 0 ? kPinned : window;
            ok = emit_program(subs, mem, cl_nodes, level, sorted_ids, prog, ps);
            if (!ok) break;
            const bool fits =
                static_cast<uint32_t>(ad_smem_layout(prog[0], prog[5], prog[22], 32).total) <=
                kAdHalfSmBytes;
            if (best_cost < 0 || (fits && !best_fits) ||
                (fits == best_fits && m_last_critical_path < best_cost)) {
              best = prog;
              best_cost = m_last_critical_path;
              best_fits = fits;
            }
            if (kPinned > 0 || legacy_schedule()) break;
          }
          if (ok) prog.swap(best);
        }
        for (int32_t nd : cl_nodes) {
          local[nd] = -1;
          main_sc.local[nd] = -1;
        }
        if (!ok) return false;
        pid = static_cast<int32_t>(ps.prog_offset.size());
        ps.prog_offset.push_back(static_cast<int64_t>(ps.blob.size()));
        ps.blob.insert(ps.blob.end(), prog.begin(), prog.end());
        const int32_t smem = static_cast<int32_t>(prog[0]) * 8;
        ps.prog_smem.push_back(smem);
        ps.prog_width.push_back(static_cast<int32_t>(prog[21]));
        ps.max_smem = std::max(ps.max_smem, smem);
        by_hash.emplace(key, pid);
      }
      if ((ps.blob.data() + ps.prog_offset[pid])[20] == 2) {
        ci.const_vals.push_back(1.0);
        ci.const_vals.push_back(-1.0);
      }
      {
        const uint32_t* P = ps.blob.data() + ps.prog_offset[pid];
        ps.n_instr += P[14];
        ps.n_visits += P[19];
        ps.n_contribs += P[18];
      }
      // --- binding ---------------------------------------------------------------
      if (ps.bindings.size() & 1) ps.bindings.push_back(0);
      ps.cluster_prog.push_back(pid);
      ps.cluster_bind.push_back(static_cast<int64_t>(ps.bindings.size()));
      auto push_i32s = [&](const std::vector<int32_t>& v) {
        const size_t off = ps.bindings.size();
        ps.bindings.resize(off + v.size());
        if (!v.empty()) std::memcpy(ps.bindings.data() + off, v.data(), v.size() * 4);
      };
      push_i32s(ci.leaf_index);
      if (ps.bindings.size() & 1) ps.bindings.push_back(0);
      {
        const size_t off = ps.bindings.size();
        ps.bindings.resize(off + ci.const_vals.size() * 2);
        if (!ci.const_vals.empty()) {
          std::memcpy(ps.bindings.data() + off, ci.const_vals.data(),
                      ci.const_vals.size() * 8);
        }
      }
      push_i32s(ci.val_out_stage);
      push_i32s(ci.adj_out_stage);
    }
    timer.lap("signatures + bindings");
    return true;
  }
};

}  // namespace

bool build_task_plan(ProgramSet& ps, int32_t smem_budget, std::string& error) {
  const int32_t n_prog = static_cast<int32_t>(ps.prog_offset.size());
  const int32_t n_clusters = static_cast<int32_t>(ps.cluster_prog.size());
  ps.task_prog.clear();
  ps.task_count.clear();
  ps.task_lanes.clear();
  ps.task_bind.clear();
  ps.task_bindings.clear();
  ps.launches.clear();
  std::vector<std::vector<int32_t>> by_prog(n_prog);
  for (int32_t c = 0; c < n_clusters; ++c) {
    by_prog[ps.cluster_prog[c]].push_back(c);
  }
  // lanes and block size of each program
  std::vector<int32_t> lanes(n_prog), threads(n_prog);
  auto task_bytes = [&](int32_t p, int32_t L) {
    const uint32_t* P = ps.blob.data() + ps.prog_offset[p];
    return ad_smem_layout(P[0], P[5], P[22], L).total;
  };
  for (int32_t p = 0; p < n_prog; ++p) {
    if (task_bytes(p, 1) > smem_budget) {
      error = "an expression cluster needs " + std::to_string(task_bytes(p, 1)) +
              " bytes of shared memory, more than one thread block has; "
              "the global-memory fallback is not implemented";
      return false;
    }
    int32_t L = 32;
    while (L > 1 && task_bytes(p, L) > smem_budget) L >>= 1;
    // no wider than the clusters available
    while (L > 1 && L / 2 >= static_cast<int32_t>(by_prog[p].size())) L >>= 1;
    lanes[p] = L;
    // one row of L threads per worker of the program's schedule (≤ 16 × 32)
    threads[p] = std::max<int32_t>(1, ps.prog_width[p]) * L;
  }
  // launches: one per block size, biggest first; inside, heavy programs first
  std::vector<int32_t> order(n_prog);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) {
    if (threads[a] != threads[b]) return threads[a] > threads[b];
    return ps.prog_smem[a] > ps.prog_smem[b];
  });
  for (int32_t p : order) {
    const uint32_t* P = ps.blob.data() + ps.prog_offset[p];
    const int32_t n_leaf = static_cast<int32_t>(P[2]);
    const int32_t n_const = static_cast<int32_t>(P[3]);
    const int32_t n_val_out = static_cast<int32_t>(P[6]);
    const int32_t n_adj_out = static_cast<int32_t>(P[7]);
    const int32_t L = lanes[p];
    const auto& list = by_prog[p];
    if (list.empty()) continue;
    if (ps.launches.empty() || ps.launches.back().threads != threads[p]) {
      ps.launches.push_back(
          {static_cast<int32_t>(ps.task_prog.size()), 0, threads[p], 0});
    }
    ProgramSet::Launch& launch = ps.launches.back();
    launch.smem_bytes = std::max(launch.smem_bytes, task_bytes(p, L));
    for (size_t k = 0; k < list.size(); k += L) {
      const int32_t cnt =
          static_cast<int32_t>(std::min<size_t>(L, list.size() - k));
      if (ps.task_bindings.size() & 1) ps.task_bindings.push_back(0);
      ps.task_prog.push_back(p);
      ps.task_count.push_back(cnt);
      ps.task_lanes.push_back(L);
      ps.task_bind.push_back(static_cast<int64_t>(ps.task_bindings.size()));
      ++launch.n_tasks;
      // per-cluster binding layout: leaf_index | pad | const_val | val_out |
      // adj_out  →  transposed per task
      const int32_t const_off_c = (n_leaf + 1) & ~1;
      auto cluster_words = [&](int32_t lane) {
        const int32_t c = list[k + (lane < cnt ? lane : 0)];
        return ps.bindings.data() + ps.cluster_bind[c];
      };
      const size_t base = ps.task_bindings.size();
      const int32_t const_off_t = (n_leaf * L + 1) & ~1;
      const size_t total = size_t(const_off_t) + size_t(2) * n_const * L +
                           size_t(n_val_out + n_adj_out) * L;
      ps.task_bindings.resize(base + total, 0);
      uint32_t* T = ps.task_bindings.data() + base;
      for (int32_t lane = 0; lane < L; ++lane) {
        const uint32_t* B = cluster_words(lane);
        for (int32_t i = 0; i < n_leaf; ++i) T[i * L + lane] = B[i];
        for (int32_t i = 0; i < n_const; ++i) {
          T[const_off_t + 2 * (i * L + lane)] = B[const_off_c + 2 * i];
          T[const_off_t + 2 * (i * L + lane) + 1] = B[const_off_c + 2 * i + 1];
        }
        const uint32_t* Bo = B + const_off_c + 2 * n_const;
        uint32_t* To = T + const_off_t + 2 * n_const * L;
        for (int32_t i = 0; i < n_val_out + n_adj_out; ++i) {
          To[i * L + lane] = Bo[i];
        }
      }
    }
  }
  return true;
}

void build_shard_plan(const ProgramSet& ps, int32_t world, ShardPlan& out) {
  out = ShardPlan{};
  out.world = world;
  std::vector<std::vector<int32_t>> slots(world);
  for (const auto& L : ps.launches) {
    for (int32_t r = 0; r < world; ++r) {
      const int32_t b = static_cast<int32_t>(int64_t(L.n_tasks) * r / world);
      const int32_t e = static_cast<int32_t>(int64_t(L.n_tasks) * (r + 1) / world);
      out.first_task.push_back(L.first_task + b);
      out.n_tasks.push_back(e - b);
      for (int32_t t = L.first_task + b; t < L.first_task + e; ++t) {
        const uint32_t* P = ps.blob.data() + ps.prog_offset[ps.task_prog[t]];
        const int32_t n_leaf = static_cast<int32_t>(P[2]);
        const int32_t n_const = static_cast<int32_t>(P[3]);
        const int32_t n_out = static_cast<int32_t>(P[6] + P[7]);
        const int32_t lanes = ps.task_lanes[t], cnt = ps.task_count[t];
        const uint32_t* B = ps.task_bindings.data() + ps.task_bind[t];
        const int32_t const_off = (n_leaf * lanes + 1) & ~1;
        const uint32_t* outs = B + const_off + 2 * n_const * lanes;
        for (int32_t i = 0; i < n_out; ++i) {
          for (int32_t c = 0; c < cnt; ++c) {
            slots[r].push_back(static_cast<int32_t>(outs[i * lanes + c]));
          }
        }
      }
    }
  }
  out.len.resize(world);
  for (int32_t r = 0; r < world; ++r) {
    out.len[r] = static_cast<int32_t>(slots[r].size());
    out.max_len = std::max(out.max_len, out.len[r]);
  }
  out.slots.assign(size_t(world) * out.max_len, -1);
  for (int32_t r = 0; r < world; ++r) {
    std::copy(slots[r].begin(), slots[r].end(),
              out.slots.begin() + size_t(r) * out.max_len);
  }
}

bool compile_autodiff(const Tape& tape, const RowSet rows[SLPB_OUT_COUNT],
                      bool ignore_h_c, CompiledAD& out,
                      const std::function<void()>& on_patterns) {
  out = CompiledAD{};
  StageTimer timer;
  Compiler C{tape, out.error};
  const int32_t n = tape.n_x, me = tape.n_y, mi = tape.n_z;

  // ---- static patterns ---------------------------------------------------------
  auto jac_pattern = [&](const RowSet& rs, int32_t nrows) {
    std::vector<std::pair<int32_t, int32_t>> posv;
    if (rs.present) {
      for (size_t k = 0; k < rs.cached_row.size(); ++k) {
        posv.emplace_back(rs.cached_col[k], rs.cached_row[k]);
      }
      for (int32_t r = 0; r < rs.n_rows; ++r) {
        if (!rs.row_swept[r]) continue;
        for (int32_t k = rs.out_ptr[r]; k < rs.out_ptr[r + 1]; ++k) {
          posv.emplace_back(rs.out_col[k], r);
        }
      }
    }
    return pattern_from_positions(nrows, n, posv);
  };
  out.A_e = jac_pattern(rows[SLPB_OUT_A_E], me);
  out.A_i = jac_pattern(rows[SLPB_OUT_A_I], mi);
  {
    std::vector<std::pair<int32_t, int32_t>> posv;
    for (int which : {SLPB_OUT_H_F, SLPB_OUT_H_C}) {
      const RowSet& rs = rows[which];
      if (!rs.present || (which == SLPB_OUT_H_C && ignore_h_c)) continue;
      for (size_t k = 0; k < rs.cached_row.size(); ++k) {
        if (rs.cached_row[k] >= rs.cached_col[k]) {
          posv.emplace_back(rs.cached_col[k], rs.cached_row[k]);
        }
      }
      for (int32_t r = 0; r < rs.n_rows; ++r) {
        if (!rs.row_swept[r]) continue;
        for (int32_t k = rs.out_ptr[r]; k < rs.out_ptr[r + 1]; ++k) {
          if (r >= rs.out_col[k]) posv.emplace_back(rs.out_col[k], r);
        }
      }
    }
    out.H = pattern_from_positions(n, n, posv);
  }
  out.off_g = 0;
  out.off_ae = n;
  out.off_ai = out.off_ae + out.A_e.nnz();
  out.off_h = out.off_ai + out.A_i.nnz();
  const int64_t n_deriv_entries = out.off_h + out.H.nnz();
  timer.lap("static patterns");
  if (on_patterns) on_patterns();
  GatherBuilder dgather{n_deriv_entries};
  GatherBuilder vgather{1 + int64_t(me) + mi};

  // scale references: −1 → 1, −2 → d_f, k ≥ 0 → [d_ce | d_ci][k]
  auto scale_of = [&](int which, int32_t row) -> int32_t {
    switch (which) {
      case SLPB_OUT_F: case SLPB_OUT_G: case SLPB_OUT_H_F: return -2;
      case SLPB_OUT_C_E: case SLPB_OUT_A_E: return row;
      case SLPB_OUT_C_I: case SLPB_OUT_A_I: return me + row;
      default: return -1;
    }
  };
  auto deriv_entry = [&](int which, int32_t row, int32_t col) -> int64_t {
    switch (which) {
      case SLPB_OUT_G: return out.off_g + col;
      case SLPB_OUT_A_E: return out.off_ae + out.A_e.find(row, col);
      case SLPB_OUT_A_I: return out.off_ai + out.A_i.find(row, col);
      default: return row >= col ? out.off_h + out.H.find(row, col) : -1;
    }
  };

  // ---- value sub-rows ------------------------------------------------------------
  // The value programs (f, c_e, c_i) share nothing with the derivative
  // programs but the tape and the row descriptors, both read-only here: they
  // are compiled on a second thread, with a compiler (scratch) of their own,
  // while this thread does the derivative set.
  std::string value_error;
  auto compile_values = [&]() -> bool {
    StageTimer vtimer;
    Compiler Cv{tape, value_error};
    std::vector<SubRow> vsubs;
    int32_t next_vstage = 0;
    auto value_entry = [&](int which, int32_t row) -> int64_t {
      return which == SLPB_OUT_F ? 0
             : which == SLPB_OUT_C_E ? 1 + row : 1 + int64_t(me) + row;
    };
    std::vector<double> vconst;
    struct PendingConst { int64_t entry; int32_t scale; double v; };
    std::vector<PendingConst> pend;
    for (int which : {SLPB_OUT_F, SLPB_OUT_C_E, SLPB_OUT_C_I}) {
      const RowSet& rs = rows[which];
      if (!rs.present) continue;
      for (int32_t r = 0; r < rs.n_rows; ++r) {
        const int32_t b = rs.row_ptr[r], e = rs.row_ptr[r + 1];
        if (e == b) {
          pend.push_back({value_entry(which, r), scale_of(which, r),
                          rs.const_val.empty() ? 0.0 : rs.const_val[r]});
          continue;
        }
        Cv.split_row(rs.row_nodes.data() + b, e - b,
                    [&](int8_t seed, std::vector<int32_t> nodes) {
                      SubRow sr;
                      sr.out = which;
                      sr.row = r;
                      sr.seed = seed;
                      sr.is_value = true;
                      sr.value_stage = next_vstage;
                      vgather.add(value_entry(which, r), next_vstage,
                                  scale_of(which, r), seed < 0);
                      ++next_vstage;
                      sr.nodes = std::move(nodes);
                      vsubs.push_back(std::move(sr));
                    });
      }
    }
    // constants go after the swept slots for the value stage
    out.value_stage_init.assign(next_vstage, 0.0);
    for (auto& p : pend) {
      vgather.add(p.entry, static_cast<int32_t>(out.value_stage_init.size()),
                  p.scale);
      out.value_stage_init.push_back(p.v);
    }
    out.value_stage_size = static_cast<int32_t>(out.value_stage_init.size());
    vtimer.lap("value sub-rows");
    if (!Cv.build_programs(vsubs, out.values)) return false;
    out.value_gather = vgather.finish();
    vtimer.lap("build_programs(values) total");
    return true;
  };
  bool values_ok = false;
  std::string values_exception;
  // (an exception must not leave the thread: std::terminate would take the
  // whole process down instead of failing this one solve)
  std::thread values_thread{[&] {
    try {
      values_ok = compile_values();
    } catch (const std::exception& e) {
      values_ok = false;
      values_exception = e.what();
    } catch (...) {
      values_ok = false;
      values_exception = "unknown exception";
    }
  }};
  struct Joiner {
    std::thread& t;
    ~Joiner() {
      if (t.joinable()) t.join();
    }
  } joiner{values_thread};

  // ---- derivative sub-rows -------------------------------------------------------
  std::vector<SubRow> dsubs;
  for (int which : {SLPB_OUT_G, SLPB_OUT_A_E, SLPB_OUT_A_I, SLPB_OUT_H_F,
                    SLPB_OUT_H_C}) {
    const RowSet& rs = rows[which];
    if (!rs.present || (which == SLPB_OUT_H_C && ignore_h_c)) continue;
    const bool hess = which == SLPB_OUT_H_F || which == SLPB_OUT_H_C;
    // constants: cached triplets of LINEAR rows
    for (size_t k = 0; k < rs.cached_row.size(); ++k) {
      const int32_t r = rs.cached_row[k], c = rs.cached_col[k];
      if (hess && r < c) continue;
      const int32_t slot = static_cast<int32_t>(out.deriv_stage_init.size());
      out.deriv_stage_init.push_back(rs.cached_val[k]);
      dgather.add(deriv_entry(which, r, c), slot, scale_of(which, r));
    }
  }
  const int32_t n_deriv_const = static_cast<int32_t>(out.deriv_stage_init.size());
  int32_t next_stage = n_deriv_const;
  for (int which : {SLPB_OUT_G, SLPB_OUT_A_E, SLPB_OUT_A_I, SLPB_OUT_H_F,
                    SLPB_OUT_H_C}) {
    const RowSet& rs = rows[which];
    if (!rs.present || (which == SLPB_OUT_H_C && ignore_h_c)) continue;
    const bool hess = which == SLPB_OUT_H_F || which == SLPB_OUT_H_C;
    for (int32_t r = 0; r < rs.n_rows; ++r) {
      if (!rs.row_swept[r]) continue;
      const int32_t b = rs.row_ptr[r], e = rs.row_ptr[r + 1];
      // outputs of the row that survive the triangle filter
      int32_t n_outs = 0;
      for (int32_t k = rs.out_ptr[r]; k < rs.out_ptr[r + 1]; ++k) {
        if (hess && r < rs.out_col[k]) continue;
        C.outcol[rs.out_node[k]] = rs.out_col[k];
        ++n_outs;
      }
      if (n_outs > 0) {
        C.split_row(rs.row_nodes.data() + b, e - b,
                    [&](int8_t seed, std::vector<int32_t> nodes) {
                      SubRow sr;
                      sr.out = which;
                      sr.row = r;
                      sr.seed = seed;
                      // outputs inside this sub-row, in list order (the order
                      // of the reference's output lists, jacobian.hpp:65-72)
                      for (int32_t nd : nodes) {
                        const int32_t col = C.outcol[nd];
                        if (col < 0) continue;
                        sr.outs.emplace_back(next_stage, nd);
                        dgather.add(deriv_entry(which, r, col), next_stage,
                                    scale_of(which, r));
                        ++next_stage;
                      }
                      if (sr.outs.empty()) return;
                      sr.nodes = std::move(nodes);
                      dsubs.push_back(std::move(sr));
                    });
      }
      for (int32_t k = rs.out_ptr[r]; k < rs.out_ptr[r + 1]; ++k) {
        C.outcol[rs.out_node[k]] = -1;
      }
    }
  }
  out.deriv_stage_size = next_stage;
  out.deriv_stage_init.resize(next_stage, 0.0);
  timer.lap("derivative sub-rows");
  if (!C.build_programs(dsubs, out.derivs)) return false;
  timer.lap("build_programs(derivs) total");
  out.deriv_gather = dgather.finish();

  values_thread.join();
  if (!values_ok) {
    out.error = values_exception.empty()
                    ? value_error
                    : "compiling the value programs: " + values_exception;
    return false;
  }
  return true;
}

}  // namespace slpb
