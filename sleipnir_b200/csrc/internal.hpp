// Internal declarations shared by the translation units of libslpb.so.
// Nothing here crosses the C ABI (include/slpb.h).
#pragma once

#include <cstdint>
#include <functional>
#include <string>
#include <utility>
#include <vector>

#include "slpb.h"

namespace slpb {

// ---------------------------------------------------------------------------
// Host copy of what the caller uploaded
// ---------------------------------------------------------------------------

struct Tape {
  int32_t n_nodes = 0;
  std::vector<uint8_t> op;
  std::vector<int32_t> lhs, rhs;
  std::vector<double> val;
  int32_t n_x = 0, n_y = 0, n_z = 0;
  /// node id → index into the concatenated leaf vector [x | d_ce⊙y | d_ci⊙z],
  /// or −1 for a node that is not a leaf of the problem.
  std::vector<int32_t> leaf_of_node;
};

struct RowSet {
  bool present = false;
  int32_t n_rows = 0, n_cols = 0;
  std::vector<int32_t> row_ptr, row_nodes, out_ptr, out_col, out_node;
  std::vector<uint8_t> row_swept;
  std::vector<int32_t> cached_row, cached_col;
  std::vector<double> cached_val;
  std::vector<double> const_val;  // value rows with an empty list
};

/// Compressed sparse column pattern (Eigen ColMajor layout, 32-bit indices).
struct Pattern {
  int32_t rows = 0, cols = 0;
  std::vector<int32_t> colptr, rowidx;
  int64_t nnz() const { return static_cast<int64_t>(rowidx.size()); }
  /// Index of entry (r, c), or −1.
  int64_t find(int32_t r, int32_t c) const;
};

// ---------------------------------------------------------------------------
// Compiled autodiff programs (built by compile.cpp, run by ad_kernels.cu)
// ---------------------------------------------------------------------------
//
// A *cluster* is a connected set of rows that share interior nodes. A
// *program* is the position-independent code of a cluster (local slot numbers
// only); clusters whose programs are identical share one copy — every time
// step of a direct transcription ends up on the same program, with its own
// *binding* (global leaf indices, constants, output slots). A *task* is up to
// 32 clusters of one program evaluated side by side by one thread block: lane =
// cluster, so every warp instruction applies one node of the graph to 32 time
// steps, and the scratch values/adjoints of the 32 clusters are interleaved in
// shared memory.
//
// Program blob, 32-bit words (every program starts 16-byte aligned):
//   [0] n_scratch    physical scratch slots per cluster (values and adjoints
//                    share them; assigned by liveness over the level schedule)
//   [1] n_logical    value slots + adjoint visits before slot sharing (stats)
//   [2] n_leaf       [3] n_const
//   [4] n_blocks     level blocks in the instruction stream
//   [5] prologue_words  header + tables: what a thread block keeps resident
//   [6] n_val_out    [7] n_adj_out
//   [8] off_leaf_slot   u16[n_leaf]   (word offsets from blob start)
//   [9] off_const_slot  u16[n_const]
//   [10] off_block_table u32[n_blocks+1]  word offset of each level block
//   [11] max_block_words  largest level block (ring-buffer stage size)
//   [12] n_fwd_levels [13] n_rev_levels [14] n_instr   (stats)
//   [15] off_val_out    u16[n_val_out]  scratch slot of each value output
//   [16] off_adj_out    u16[n_adj_out]  scratch slot of each derivative output
//   [17] max_width      widest level (forward or reverse)
//   [18] n_contrib [19] n_visits (stats)
//   [20] n_synth     trailing constants +1, −1 appended by the compiler
//   [21] n_workers   rows of lanes that run private item lists (≤ 16)
//   [22] ring_words  words of shared memory the instruction stream needs
//   [23] resident    1: the whole stream is copied in once (it is small);
//                    0: kAdStages blocks travel through a ring
//   tables …, then the INSTRUCTION STREAM: one 16-byte-aligned block per
//   SUPER-LEVEL, in execution order (the thread block synchronises between
//   blocks only): {kind, n_items, n_contrib, W}, then for W > 0 the W + 1 item
//   offsets of the workers' private lists (padded to an even word count), then
//     kBlockForward:  FwdInstr[n_items]      worker w runs [off[w], off[w+1]) in order
//     kBlockValueOut: nothing (value outputs are stored at this point)
//     kBlockReverse:  Visit[n_items], Contrib[n_contrib]
//   The device streams these blocks through a shared-memory ring with TMA bulk
//   copies (cp.async.bulk + mbarrier) a few levels ahead of their use.
// Binding record of a cluster, 32-bit words:
//   leaf_index  i32[n_leaf]      index into the leaf vector
//   const_val   f64[n_const]     (8-byte aligned)
//   val_out     i32[n_val_out]   stage slot of each value output
//   adj_out     i32[n_adj_out]   stage slot of each derivative output
// Binding block of a task: the same four arrays, each transposed to
// [entry][lanes] so that the lanes of a warp read consecutive words.

constexpr int kProgHeaderWords = 24;
constexpr uint32_t kBlockForward = 0, kBlockReverse = 1, kBlockValueOut = 2;
/// Contrib.op of a contribution that is adjoint × (value in slot l): the
/// partials of +, −, unary − (multiplier ±1) and × (the other operand).
constexpr uint8_t kOpLinear = 255;
/// The same with the product subtracted: the partial of cos(x) is
/// a·(−sin x) = −(a·sin x) exactly, taken from the slot of a sin(x) node of the
/// same cluster instead of evaluating the sine again.
constexpr uint8_t kOpLinearNeg = 254;
constexpr int kAdStages = 3;  // ring-buffer stages of the instruction stream
/// An instruction stream stays resident in shared memory when it is at most
/// this big AND a full task (32 lanes of scratch + stream) still leaves room
/// for a second task on the SM; otherwise it travels through the ring.
constexpr uint32_t kAdResidentBytes = 64 * 1024;
constexpr uint32_t kAdHalfSmBytes = 112 * 1024;

/// Shared-memory layout of one task (bytes from the start of dynamic shared
/// memory): scratch | header+tables | instruction ring | mbarriers.
struct AdSmemLayout {
  int32_t off_prologue, off_ring, off_bars, total;
};
#if defined(__CUDACC__)
__host__ __device__
#endif
inline AdSmemLayout ad_smem_layout(uint32_t n_scratch, uint32_t prologue_words,
                                   uint32_t ring_words, int lanes) {
  AdSmemLayout L;
  const uint32_t scratch = (n_scratch * uint32_t(lanes) * 8u + 15u) & ~15u;
  L.off_prologue = static_cast<int32_t>(scratch);
  L.off_ring =
      static_cast<int32_t>((scratch + prologue_words * 4u + 15u) & ~15u);
  L.off_bars = L.off_ring + static_cast<int32_t>(ring_words) * 4;
  L.total = L.off_bars + kAdStages * 8;
  return L;
}

struct FwdInstr {  // 8 bytes
  uint16_t dst, a, b;
  uint8_t op, pad;
};
struct Visit {  // 8 bytes
  uint32_t contrib_begin;  // index into the program's Contrib array
  uint16_t adj;            // adjoint slot written by this visit
  uint8_t n_contrib;       // 255 = more than 254: count in next visit's begin
  int8_t seed;             // root visits: adjoint = seed (±1); else 0
};
struct Contrib {  // 8 bytes
  uint16_t parent_adj, l, r;
  uint8_t op, side;  // side: 0 → ∂/∂lhs, 1 → ∂/∂rhs
};

struct ProgramSet {
  // one entry per distinct program
  std::vector<uint32_t> blob;          // all programs back to back
  std::vector<int64_t> prog_offset;    // word offset of each program
  std::vector<int32_t> prog_smem;      // bytes of shared memory per cluster
  std::vector<int32_t> prog_width;     // workers of the program's schedule
  // one entry per cluster
  std::vector<int32_t> cluster_prog;
  std::vector<int64_t> cluster_bind;   // word offset into `bindings`
  std::vector<uint32_t> bindings;
  int64_t n_instr = 0, n_visits = 0, n_contribs = 0;  // totals over clusters
  int32_t max_smem = 0;                // largest per-cluster scratch, bytes

  // ---- task plan (build_task_plan) ----
  std::vector<int32_t> task_prog, task_count, task_lanes;
  std::vector<int64_t> task_bind;      // word offset into task_bindings
  std::vector<uint32_t> task_bindings;
  struct Launch {
    int32_t first_task, n_tasks, threads, smem_bytes;
  };
  std::vector<Launch> launches;        // tasks are sorted by launch
};

/// Multi-GPU sharding of one program set: every launch's tasks are split into
/// `world` contiguous ranges; rank r evaluates range r and the ranks exchange
/// the stage slots they produced (one all-gather of `max_len` doubles per rank).
struct ShardPlan {
  int32_t world = 1;
  /// per launch and rank: first task and task count (launch-major)
  std::vector<int32_t> first_task, n_tasks;
  /// per rank: the stage slots its tasks write, padded with −1 to max_len
  std::vector<int32_t> slots;  // world × max_len
  std::vector<int32_t> len;    // world
  int32_t max_len = 0;
};
void build_shard_plan(const ProgramSet& ps, int32_t world, ShardPlan& out);

/// Groups the clusters of `ps` into tasks and launches. smem_budget: bytes of
/// shared memory one task may use. Returns false (error set) when a single
/// cluster does not fit.
bool build_task_plan(ProgramSet& ps, int32_t smem_budget, std::string& error);

/// out[e] = Σ_k scale(k) · stage[src_idx[k]] over k ∈ [ptr[e], ptr[e+1]).
struct Gather {
  std::vector<int32_t> ptr;        // n_entries + 1
  std::vector<int32_t> src_idx;    // stage slot
  std::vector<int32_t> src_scale;  // −1: 1, −2: d_f, ≥0: index into [d_ce|d_ci]
  int32_t n_entries() const { return static_cast<int32_t>(ptr.size()) - 1; }
};

struct CompiledAD {
  // value programs (f, c_e, c_i) and derivative programs (g, A_e, A_i, H)
  ProgramSet values, derivs;
  // stage arrays: [constants | swept outputs]; constants are uploaded once
  std::vector<double> value_stage_init, deriv_stage_init;
  int32_t value_stage_size = 0, deriv_stage_size = 0;
  // final entries, in this order: values: [f | c_e | c_i];
  // derivatives: [g (n) | A_e.val | A_i.val | H.val]
  Gather value_gather, deriv_gather;
  Pattern A_e, A_i, H;  // H: lower triangle of d_f·H_f + H_c
  int64_t off_g = 0, off_ae = 0, off_ai = 0, off_h = 0;  // into deriv entries
  std::string error;
};

bool ingest_tape(Tape& t, int32_t n_nodes, const uint8_t* op,
                 const int32_t* lhs, const int32_t* rhs, const double* val,
                 int32_t n_x, const int32_t* leaf_x, int32_t n_y,
                 const int32_t* leaf_y, int32_t n_z, const int32_t* leaf_z,
                 std::string& error);
bool ingest_rows(RowSet& r, const Tape& t, int which, const slpb_rowset* rows,
                 const double* const_val, std::string& error);

/// on_patterns (optional) is called as soon as out.A_e, out.A_i and out.H — the
/// static sparsity patterns — are final, while the programs are still being
/// built (the caller may start the symbolic analysis of the KKT system then).
bool compile_autodiff(const Tape& tape, const RowSet rows[SLPB_OUT_COUNT],
                      bool ignore_h_c, CompiledAD& out,
                      const std::function<void()>& on_patterns = {});

// ---------------------------------------------------------------------------
// Symbolic analysis of the reduced KKT system (symbolic.cpp)
// ---------------------------------------------------------------------------

/// How each lower-triangle entry of lhs = [H + tril(A_iᵀΣA_i); A_e] is
/// computed from H.val, A_e.val, A_i.val and Σ.
struct KktRecipe {
  Pattern K;                       // lower triangle incl. full diagonal
  std::vector<int32_t> h_idx;      // per K entry: index into H.val or −1
  std::vector<int32_t> ae_idx;     // per K entry: index into A_e.val or −1
  std::vector<int32_t> prod_ptr;   // per K entry: range of A_iᵀΣA_i terms
  std::vector<int32_t> prod_a, prod_b, prod_row;  // A_i.val idx ×2, ineq row
  std::vector<int32_t> diag_idx;   // per column: K entry of the diagonal
};

void build_kkt_recipe(int32_t n, int32_t me, const Pattern& H,
                      const Pattern& A_e, const Pattern& A_i, KktRecipe& out);

/// Supernodal multifrontal plan. Fronts are dense, column-major, of order
/// `front_dim`, the first `n_piv` rows/cols being the supernode's own columns.
struct Symbolic {
  int32_t dim = 0;
  std::vector<int32_t> perm, iperm;  // perm[k] = original index eliminated k-th
  std::vector<int32_t> parent;       // column elimination tree (permuted)
  int64_t nnz_l = 0;
  int32_t etree_height = 0;

  int32_t n_super = 0;
  std::vector<int32_t> super_first;  // n_super+1: column ranges (permuted)
  std::vector<int32_t> super_parent; // assembly tree
  std::vector<int32_t> super_level;  // 0 = leaves
  std::vector<int32_t> level_ptr, level_supers;  // supernodes grouped by level
  std::vector<int32_t> front_dim;    // order of each front
  std::vector<int64_t> rows_ptr;     // n_super+1 → front row indices (permuted)
  std::vector<int32_t> rows_idx;
  std::vector<int64_t> panel_ptr;    // n_super+1 → L panel storage (dim×n_piv)
  std::vector<int64_t> update_ptr;   // n_super+1 → update matrix storage
  std::vector<int64_t> child_ptr;    // n_super+1 → children lists
  std::vector<int32_t> child_idx;
  std::vector<int64_t> rel_ptr;      // per supernode: where its update rows
  std::vector<int32_t> rel_idx;      //   land in the parent's front
  // scatter of K values into fronts: per supernode a list of
  // (K entry, position r + c·front_dim)
  std::vector<int64_t> asm_ptr;
  std::vector<int32_t> asm_src, asm_dst;
  std::vector<uint8_t> col_is_primal;  // permuted column < n in original order
  int32_t max_front = 0, n_levels = 0;
  int64_t panel_size = 0, update_size = 0;
};

bool analyze_kkt(const Pattern& K, int32_t n_primal, int ordering,
                 const int32_t* user_perm, Symbolic& out, std::string& error);

/// Multi-GPU partition of the assembly tree: the top of the tree
/// (the interface between the time chunks) is replicated on every rank, every
/// subtree hanging below it is owned by one rank. A rank eliminates its own
/// subtrees locally, the ranks exchange the subtree roots' update matrices,
/// update vectors and inertia counts in one small all-gather, every rank then
/// eliminates the top redundantly and back-substitutes into its own subtrees.
struct TreeShard {
  int32_t world = 1;
  std::vector<int32_t> owner;        // per front: rank, or −1 = top
  std::vector<int32_t> top_order;    // top fronts by ascending level
  /// per rank: own fronts by ascending level, and the roots of its subtrees
  std::vector<std::vector<int32_t>> rank_order, rank_roots;
  /// per front (top fronts only, 0 elsewhere): children that are NOT top —
  /// complete by the time the top is processed
  std::vector<int32_t> top_fcount_init;
  std::vector<double> rank_work;     // Σ np·F² of what each rank owns
  double top_work = 0.0;
};
void build_tree_shard(const Symbolic& Y, int32_t world, TreeShard& out);

std::vector<int32_t> order_nested_dissection(const Pattern& lowerK);
std::vector<int32_t> order_amd(const Pattern& lowerK);

}  // namespace slpb
