// libslpb.so — CUDA kernels (sm_100a) and the C ABI of include/slpb.h.
//
// Data layout in HBM (all FP64 unless noted), resident for a whole solve:
//   iterate      x[n] s[mi] y[me] z[mi]      + trial copies
//   leaves       [x | d_ce⊙y | d_ci⊙z]       what the tape's VAR nodes read
//   programs     cluster programs + bindings (u32 words; see internal.hpp)
//   stage        [constants | swept outputs] per program set
//   values       [f | c_e | c_i]             current and trial
//   derivatives  [g | A_e.val | A_i.val | H.val]   (CSC patterns are static)
//   KKT          Kval[nnz(lhs)] in the reference's column-major lower order
//   factor       supernodal panels, update matrices, D, in elimination order
// Every kernel below is launched on the handle's stream; the host only reads
// back small result structs through pinned memory.
//
// There is deliberately no CPU implementation behind this ABI: without a CUDA
// device slpb_create fails with SLPB_ERR_NO_DEVICE.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <future>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#ifdef SLPB_SWEEP_STAMPS
// development build (make EXTRA=-DSLPB_SWEEP_STAMPS; scripts/sweep_debug.py):
// block 0 of a 512-thread sweep (16 workers × 32 lanes) records clock64() at
// its phase boundaries
__device__ long long g_sweep_stamps[128];  // block 0 | block 75 (alone on its SM)
#ifdef __CUDA_ARCH__
#define SLPB_AD_STAMP(k)                                                      \
  do {                                                                        \
    if (threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == 75) &&         \
        blockDim.x == 512 && (k) < 64)                                        \
      g_sweep_stamps[(blockIdx.x ? 64 : 0) + (k)] = clock64();                \
  } while (0)
#else
#define SLPB_AD_STAMP(k)
#endif
#endif
#include "ad_core.hpp"
#include "internal.hpp"
#include "kkt_core.hpp"
#include "ldlt_core.hpp"
#include "ldlt_warp.cuh"
#include "slpb.h"

namespace slpb {

// ---------------------------------------------------------------------------
// NCCL, resolved at run time (dlopen): only the multi-GPU sharded sweep needs
// it, and a process that already carries an NCCL (torch) shares that copy.
// Minimal declarations of the stable NCCL 2 C API.
// ---------------------------------------------------------------------------
struct NcclUniqueId {
  char internal[128];
};
using NcclComm = void*;
struct NcclApi {
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) =
      nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};
constexpr int kNcclFloat64 = 8;  // ncclDouble

inline const NcclApi& nccl_api() {
  static const NcclApi api = [] {
    NcclApi a;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return a;
    a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(
        dlsym(h, "ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(
        dlsym(h, "ncclCommInitRank"));
    a.AllGather =
        reinterpret_cast<decltype(a.AllGather)>(dlsym(h, "ncclAllGather"));
    a.CommDestroy =
        reinterpret_cast<decltype(a.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(
        dlsym(h, "ncclGetErrorString"));
    a.ok = a.GetUniqueId && a.CommInitRank && a.AllGather && a.CommDestroy;
    return a;
  }();
  return api;
}

// ---------------------------------------------------------------------------
// small RAII helpers
// ---------------------------------------------------------------------------

/// Stream on which the DevBufs of the running ABI call allocate and free
/// (stream-ordered allocator: cudaMallocAsync / cudaFreeAsync). A plain
/// cudaFree synchronises the whole device and takes the driver's global lock —
/// ~130 of them made the teardown of one solve take anywhere from 10 ms to
/// seconds; stream-ordered frees are queued and the memory stays in the
/// device's pool for the next solve.
inline thread_local cudaStream_t tls_alloc_stream = nullptr;
struct AllocScope {
  cudaStream_t prev;
  explicit AllocScope(cudaStream_t st) : prev{tls_alloc_stream} {
    tls_alloc_stream = st;
  }
  ~AllocScope() { tls_alloc_stream = prev; }
};

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  cudaStream_t owner = nullptr;  // stream the allocation is ordered on
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p) {
      // free on the stream of the running call if there is one (it may differ
      // from `owner` only by being the same handle's stream), else the owner's
      cudaStream_t st = tls_alloc_stream ? tls_alloc_stream : owner;
      if (cudaFreeAsync(p, st) != cudaSuccess) {
        cudaGetLastError();
        cudaFree(p);
      }
    }
    p = nullptr;
    n = 0;
  }
  cudaError_t alloc(size_t count) {
    release();
    n = count;
    if (count == 0) return cudaSuccess;
    owner = tls_alloc_stream;
    return cudaMallocAsync(reinterpret_cast<void**>(&p), count * sizeof(T),
                           owner);
  }
  cudaError_t upload(const std::vector<T>& v, cudaStream_t st) {
    cudaError_t e = alloc(v.size());
    if (e != cudaSuccess || v.empty()) return e;
    return cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T),
                           cudaMemcpyHostToDevice, st);
  }
  cudaError_t zero(cudaStream_t st) {
    if (!n) return cudaSuccess;
    return cudaMemsetAsync(p, 0, n * sizeof(T), st);
  }
};

struct DevProgramSet {
  DevBuf<uint32_t> blob, task_bindings;
  DevBuf<int64_t> prog_offset, task_bind;
  DevBuf<int32_t> task_prog, task_count, task_lanes;
  std::vector<ProgramSet::Launch> launches;
  int n_tasks = 0, max_smem = 0;
};

struct DevGather {
  DevBuf<int32_t> ptr, src_idx, src_scale;
  DevBuf<int32_t> long_entries;  // entries with many sources
  int n_entries = 0, n_long = 0;
};

constexpr int kResultDoubles = 64;
constexpr int kLongGather = 48;  // sources above which a block reduces an entry
constexpr int kReduceThreads = 256;
constexpr int kReduceBlocks = 296;  // two per SM

}  // namespace slpb

using namespace slpb;

struct slpb_group;
struct slpb_solver {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string error = "";
  // host side
  Tape tape;
  RowSet rows[SLPB_OUT_COUNT];
  CompiledAD ad;
  KktRecipe recipe;
  Symbolic sym;
  bool have_tape = false, finalized = false, analyzed = false;
  bool ignore_h_c = false;
  int n = 0, me = 0, mi = 0, dim = 0;
  double d_f = 1.0;
  slpb_counters counters{};
  // device: programs
  DevProgramSet pv, pd;
  DevGather gv, gd;
  DevBuf<double> vstage, dstage;
  // device: state
  DevBuf<double> leaf_cur, leaf_trial, d_c, inv_d_c;  // inv_d_c = RN(1 / d_c)
  DevBuf<double> x, s, y, z, tx, ts, ty, tz;
  DevBuf<double> vals_cur, vals_trial, dvals;
  DevBuf<int32_t> ae_colptr, ae_rowidx, ai_colptr, ai_rowidx;
  DevBuf<int32_t> ai_rowptr, ai_rcol, ai_ridx;  // A_i by rows
  // device: KKT
  DevBuf<int32_t> k_h_idx, k_ae_idx, k_prod_ptr, k_prod_a, k_prod_b, k_prod_row;
  DevBuf<uint8_t> k_diag_flag;  // KKT entry is a primal diagonal entry
  DevBuf<double> Kval, sigma, sinv, tvec, rhs, sol;
  // device: symbolic + factor
  DevBuf<int32_t> sy_super_first, sy_front_dim, sy_rows_idx, sy_child_idx,
      sy_rel_idx, sy_asm_src, sy_asm_dst, sy_asm_dst_ld, sy_ext_src, sy_ext_dst, sy_perm,
      sy_level_supers;
  DevBuf<int64_t> sy_rows_ptr, sy_panel_ptr, sy_update_ptr, sy_child_ptr,
      sy_rel_ptr, sy_asm_ptr;
  DevBuf<uint8_t> sy_col_is_primal;
  DevBuf<double> panels, updates, D, uvecs, xperm;
  DevBuf<int32_t> fstats;  // FactorStats as 6 ints
  DevBuf<int32_t> sy_super_parent, sy_nchild, tree_sync;
  // Fronts above order 32 (a dense row of the KKT system, e.g. a time-step
  // variable shared by every stage) and their ancestors — the "top" — take one
  // block per front and one launch per level; everything below them, nearly
  // all of the tree, still takes the warp-per-front tree kernels.
  bool hybrid = false;
  int n_small = 0;
  // fronts beyond the shared-memory cap of one block: global workspaces
  DevBuf<double> front_scratch;
  DevBuf<int32_t> front_slot;
  int64_t front_stride = 0;
  int front_cap_doubles = 0, front_smem_bytes = 0;
  DevBuf<int32_t> hy_small_order, hy_top_supers, hy_bflag_init;
  std::vector<int32_t> hy_level_off;  // offsets of the top's levels in hy_top_supers
  DevBuf<int32_t> solve_sync;       // dependency words of k_solve_tree
  bool solve_sync_preset = false;   // written by the last factor launch's init
  DevBuf<FrontMeta> sy_metas;
  DevBuf<unsigned long long> tree_debug;
  bool use_tree = false;
  int factor_arith = SLPB_ARITH_REFERENCE;  // slpb_set_factor_arithmetic
  std::vector<int32_t> ext_begin, ext_chunks;  // per front (analysis scratch)
  Symbolic sym_ahead;        // default-order analysis started inside slpb_finalize
  bool sym_ahead_ok = false;
  int tree_blocks = 0, solve_blocks = 0, tree_smem_doubles = 0;
  int factor_sel = 0;  // which variant of the last factorisation the solves use
  // forward substitution fused into the factorisation (slpb_prepare_rhs)
  bool rhs_ready = false;
  double rhs_mu = 0.0;
  bool fwd_valid[2] = {false, false};
  SymbolicView sview{};
  // device: steps
  DevBuf<double> px, ps, py, pz, spx, sps, spy, spz, ce_soc, cis_soc;
  // results
  DevBuf<double> red_partials;
  DevBuf<unsigned int> red_counter;
  DevBuf<double> d_results;
  DevBuf<unsigned char> l2_flush;
  double* h_results = nullptr;  // pinned
  // multi-GPU sharded re-linearisation
  int rank = 0, world = 1;
  NcclComm comm = nullptr;
  ShardPlan shard;
  DevBuf<int32_t> shard_slots;
  DevBuf<double> shard_buf, agree_buf;
  // multi-GPU sharded factorisation / solve (TreeShard)
  bool tree_sharded = false;
  TreeShard tshard;
  DevBuf<int32_t> ts_my_order, ts_top_order, ts_bwd_order, ts_top_fcount_init;
  DevBuf<int64_t> ts_pack_pos;   // world × ts_pack_max
  DevBuf<int32_t> ts_pack_len, ts_sol_idx, ts_sol_len;
  DevBuf<double> ts_pack_buf, ts_sol_buf;
  int ts_pack_max = 0, ts_sol_max = 0;
  std::vector<int32_t> ts_pack_len_host, ts_sol_len_host;
  slpb_comm_stats comm_stats{};
  cudaEvent_t cev[6] = {};
  bool cpending[3] = {false, false, false};
  int64_t comm_timed[3] = {0, 0, 0};  // calls that were timed
  bool comm_warm[3] = {false, false, false};  // first sample of a kind dropped
  // dynamic batching with other solvers of the same pattern (group.cuh)
  slpb_group* group = nullptr;
  int group_slot = -1;
  // timing
  cudaEvent_t ev[10] = {};
  // side streams for the launches of one sweep (different program classes are
  // independent of each other): fork/join around run_sweep
  bool serial_sweeps = std::getenv("SLPB_SERIAL_SWEEPS") != nullptr;
  // per-phase device timers (CUDA events around the kernel groups;
  // slpb_get_timers). Timing events are not free — ten records per Newton
  // iteration cost 9 % of the step (gaps between back-to-back kernels) — so
  // every group is SAMPLED: one launch in timer_every carries events, the
  // averages (total_ms / count) are over the sampled launches.
  int timer_every = [] {
    if (std::getenv("SLPB_NO_TIMERS")) return 0;
    const char* e = std::getenv("SLPB_TIMER_EVERY");
    return e ? std::max(1, std::atoi(e)) : 8;
  }();
  unsigned timer_tick[5] = {0, 0, 0, 0, 0};
  bool timing_now[5] = {false, false, false, false, false};
  static constexpr int kSideStreams = 3;
  cudaStream_t side[kSideStreams] = {};
  cudaEvent_t fork_ev = nullptr, join_ev[kSideStreams] = {};
  float last_ms[5] = {0, 0, 0, 0, 0};
  slpb_timers timers{};
  bool pending[5] = {false, false, false, false, false};
};

namespace slpb {

#define CU(call)                                                          \
  do {                                                                    \
    cudaError_t e_ = (call);                                              \
    if (e_ != cudaSuccess) {                                              \
      S->error = std::string(#call) + ": " + cudaGetErrorString(e_);      \
      return SLPB_ERR_CUDA;                                               \
    }                                                                     \
  } while (0)

void harvest_comm_timers(slpb_solver* S);
// group.cuh: the member's request joins the group's next batched launch
int group_factor(slpb_solver* S, int nv, const double* delta,
                 const double* gamma, slpb_factor_info* info);
int group_solve(slpb_solver* S);
int group_download_d(slpb_solver* S, double* dst);

inline int fail(slpb_solver* S, int code, const std::string& msg) {
  S->error = msg;
  return code;
}

// ---------------------------------------------------------------------------
// kernels: autodiff
// ---------------------------------------------------------------------------

struct WarpSync {
  __device__ __forceinline__ void operator()() const { __syncwarp(); }
};
struct BlockSync {
  __device__ __forceinline__ void operator()() const { __syncthreads(); }
};

struct WarpSyncMask {
  unsigned mask;
  __device__ __forceinline__ void operator()() const { __syncwarp(mask); }
};

struct AdTasks {
  const uint32_t* blob;
  const int64_t* prog_offset;
  const int32_t* task_prog;
  const int32_t* task_count;
  const int32_t* task_lanes;
  const int64_t* task_bind;
  const uint32_t* task_bindings;
};

// ---- TMA-staged instruction stream ------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(count)
               : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
/// 1-D TMA bulk copy global → shared, completion counted on an mbarrier.
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src,
                                             uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

/// The level blocks of a program travel global → shared memory through a ring
/// of kAdStages buffers, kAdStages − 1 levels ahead of their use: thread 0
/// issues one TMA bulk copy per block, every thread waits on the stage's
/// mbarrier before decoding it, and the block's level barrier frees the stage.
struct TmaStream {
  const uint32_t* Pg;      // program in global memory
  const uint32_t* table;   // block table (shared memory copy)
  uint32_t* ring;
  uint64_t* bars;
  int n_blocks, stage_words;
  bool issuer;
  bool resident;           // the whole stream sits in `ring`, copied once
  __device__ __forceinline__ void issue(int b) const {
    const int st = b % kAdStages;
    const uint32_t bytes = (table[b + 1] - table[b]) * 4u;
    mbar_expect_tx(&bars[st], bytes);
    tma_bulk_g2s(ring + st * stage_words, Pg + table[b], bytes, &bars[st]);
  }
  /// Called by the issuing thread once, before the first acquire.
  __device__ __forceinline__ void start() const {
    if (resident) {
      // small programs (a direct-transcription stage is ≈30 KB): ONE phase of
      // bulk copies brings every block in; nothing is waited for afterwards
      const uint32_t total = (table[n_blocks] - table[0]) * 4u;
      mbar_expect_tx(&bars[0], total);
      for (uint32_t done = 0; done < total; done += 32768u) {
        const uint32_t bytes = min(32768u, total - done);
        tma_bulk_g2s(reinterpret_cast<unsigned char*>(ring) + done,
                     reinterpret_cast<const unsigned char*>(Pg + table[0]) + done,
                     bytes, &bars[0]);
      }
    } else {
      for (int b = 0; b < kAdStages && b < n_blocks; ++b) issue(b);
    }
  }
  __device__ __forceinline__ const uint32_t* acquire(int b) const {
    if (resident) {
      mbar_wait(&bars[0], 0);
      return ring + (table[b] - table[0]);
    }
    const int st = b % kAdStages;
    mbar_wait(&bars[st], (b / kAdStages) & 1);
    return ring + st * stage_words;
  }
  /// Called by every thread after the level barrier of block b.
  __device__ __forceinline__ void release(int b) const {
    if (!resident && issuer && b + kAdStages < n_blocks) issue(b + kAdStages);
  }
};

/// One thread block per task = up to 32 clusters (time steps) that share one
/// program. Lane = cluster: a warp applies ONE graph node to 32 time steps, so
/// the opcode is warp-uniform and the interleaved scratch (values and adjoints
/// of all 32 clusters) in shared memory is read without bank conflicts; the
/// warps of the block take the nodes of a level in turn and the block
/// synchronises between levels. The program itself is streamed through shared
/// memory by TMA (TmaStream). Replaces update_values + append_triplets
/// (expression_graph.hpp:85-153) for all rows of the clusters.
__global__ void __launch_bounds__(512)
k_ad_sweep(AdTasks A, int first_task, const double* __restrict__ leaf,
           double* __restrict__ stage) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SLPB_AD_STAMP(0);
  const int t = first_task + blockIdx.x;
  const uint32_t* P = A.blob + A.prog_offset[A.task_prog[t]];
  const uint32_t* B = A.task_bindings + A.task_bind[t];
  const int count = A.task_count[t];
  const int lanes = A.task_lanes[t];
  const int tid = threadIdx.x, nt = blockDim.x;
  const uint32_t n_scratch = __ldg(P + 0), pro_words = __ldg(P + 5);
  const uint32_t stage_words = __ldg(P + 11);
  const AdSmemLayout lay = ad_smem_layout(n_scratch, pro_words, __ldg(P + 22), lanes);
  double* scratch = reinterpret_cast<double*>(smem_raw);
  uint32_t* H = reinterpret_cast<uint32_t*>(smem_raw + lay.off_prologue);
  uint32_t* ring = reinterpret_cast<uint32_t*>(smem_raw + lay.off_ring);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + lay.off_bars);
  // header + tables → shared memory (16-byte vectors)
  {
    const uint4* src = reinterpret_cast<const uint4*>(P);
    uint4* dst = reinterpret_cast<uint4*>(H);
    for (uint32_t i = tid; i < pro_words / 4; i += nt) dst[i] = __ldg(src + i);
  }
  if (tid == 0) {
    for (int s = 0; s < kAdStages; ++s) mbar_init(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  TmaStream stream{P, H + H[10], ring, bars, static_cast<int>(H[4]),
                   static_cast<int>(stage_words), tid == 0, H[23] != 0};
  if (tid == 0) stream.start();
  SLPB_AD_STAMP(1);
  const BlockSync sync{};
  switch (lanes) {
    case 32: ad_run_group<32>(tid, nt, count, H, stream, B, leaf, stage, scratch, sync); break;
    case 16: ad_run_group<16>(tid, nt, count, H, stream, B, leaf, stage, scratch, sync); break;
    case 8: ad_run_group<8>(tid, nt, count, H, stream, B, leaf, stage, scratch, sync); break;
    case 4: ad_run_group<4>(tid, nt, count, H, stream, B, leaf, stage, scratch, sync); break;
    case 2: ad_run_group<2>(tid, nt, count, H, stream, B, leaf, stage, scratch, sync); break;
    default: ad_run_group<1>(tid, nt, count, H, stream, B, leaf, stage, scratch, sync); break;
  }
}

__global__ void k_gather(const int32_t* __restrict__ ptr,
                         const int32_t* __restrict__ src_idx,
                         const int32_t* __restrict__ src_scale,
                         const double* __restrict__ stage, double d_f,
                         const double* __restrict__ d_c,
                         double* __restrict__ out, int n_entries) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_entries) return;
  if (ptr[e + 1] - ptr[e] > kLongGather) return;  // k_gather_long does these
  out[e] = gather_entry(e, ptr, src_idx, src_scale, stage, d_f, d_c);
}

/// One block per entry with a long source list (a split Σ_k cost): fixed-shape
/// tree reduction, deterministic from run to run.
__global__ void __launch_bounds__(1024)
k_gather_long(const int32_t* __restrict__ entries,
              const int32_t* __restrict__ ptr,
              const int32_t* __restrict__ src_idx,
              const int32_t* __restrict__ src_scale,
              const double* __restrict__ stage, double d_f,
              const double* __restrict__ d_c, double* __restrict__ out) {
  __shared__ double red[1024];
  const int e = entries[blockIdx.x];
  const int b = ptr[e], en = ptr[e + 1];
  // one scale for the whole entry (the usual case: a split Σ_k cost) is
  // applied once to the sum, as gather_entry does per run
  const int32_t sc0 = src_scale[b];
  const bool uniform = src_scale[en - 1] == sc0;
  auto term = [&](int k) -> double {
    const int32_t raw = src_idx[k];
    double v = stage[raw & 0x7fffffff];
    if (raw < 0) v = -v;
    if (uniform) return v;
    const int32_t sc = src_scale[k];
    if (sc == -2) {
      v = d_f * v;
    } else if (sc >= 0) {
      v = d_c[sc] * v;
    }
    return v;
  };
  // fixed assignment k ≡ tid (mod blockDim): deterministic from run to run
  double acc = 0.0;
  for (int k = b + threadIdx.x; k < en; k += blockDim.x) acc += term(k);
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int w = blockDim.x / 2; w > 0; w >>= 1) {
    if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0 && uniform) {
    if (sc0 == -2) {
      red[0] = d_f * red[0];
    } else if (sc0 >= 0) {
      red[0] = d_c[sc0] * red[0];
    }
  }
  if (threadIdx.x == 0) out[e] = red[0];
}

/// Sharded sweep: buf[i] = stage[slots[i]] for the slots this rank produced.
__global__ void k_shard_pack(const double* __restrict__ stage,
                             const int32_t* __restrict__ slots, int len,
                             double* __restrict__ buf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < len) buf[i] = stage[slots[i]];
}
/// After the all-gather: stage[slots[r][i]] = buf[r][i] for every other rank.
__global__ void k_shard_unpack(const double* __restrict__ buf,
                               const int32_t* __restrict__ slots, int max_len,
                               int world, int rank, double* __restrict__ stage) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= int64_t(world) * max_len) return;
  if (i / max_len == rank) return;
  const int32_t slot = slots[i];
  if (slot >= 0) stage[slot] = buf[i];
}

__global__ void k_prepare_leaves(const double* __restrict__ x,
                                 const double* __restrict__ y,
                                 const double* __restrict__ z,
                                 const double* __restrict__ d_c, int n, int me,
                                 int mi, double* __restrict__ leaf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    leaf[i] = x[i];
  } else if (i < n + me) {
    leaf[i] = d_c[i - n] * y[i - n];
  } else if (i < n + me + mi) {
    leaf[i] = d_c[i - n] * z[i - n - me];
  }
}

// ---------------------------------------------------------------------------
// kernels: reductions. Grid-wide, two-phase and deterministic: every block
// reduces its grid-stride slice in a fixed tree, writes one partial per
// quantity, and the block that takes the last ticket combines the partials in
// a fixed order and finalises the result (no floating-point atomics anywhere).
// ---------------------------------------------------------------------------

enum RedOp { RED_SUM = 0, RED_MAX = 1, RED_MIN = 2 };

__device__ __forceinline__ double red_combine(int op, double a, double b) {
  return op == RED_SUM ? a + b : (op == RED_MAX ? fmax(a, b) : fmin(a, b));
}
__device__ __forceinline__ double red_identity(int op) {
  return op == RED_SUM ? 0.0 : (op == RED_MAX ? -INFINITY : INFINITY);
}

/// Many quantities (k_kkt_stats: 29): the warp-level tree runs through SHARED
/// MEMORY, eight quantities at a time: lane l adds the value of lane l + o for
/// o = 16, 8, …, 1 — the pairing of a __shfl_down tree, hence the same bits —
/// because 64-bit warp shuffles are the bottleneck there (2 × 29 × 5 shuffles
/// per warp, ≈4 cycles each per SM: 4.7 µs per block, measured with clock64()
/// stamps; 27.6 → 25.1 µs per launch). Few quantities keep the shuffles.
template <int NV>
__device__ __forceinline__ void block_reduce(double (&v)[NV], const int (&op)[NV],
                                             double* out) {
  constexpr int kChunk = 8;
  __shared__ double tile[kChunk][kReduceThreads];
  __shared__ double red[NV][kReduceThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  auto warp_tree = [&](double (&a)[NV]) {
    // a[q] of lane 0 ends up holding the warp's result
    if constexpr (NV <= 16) {
      // few quantities: shuffles, step by step over all of them (the NV
      // shuffles of a step are independent and overlap their latency)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        double b[NV];
#pragma unroll
        for (int q = 0; q < NV; ++q) b[q] = __shfl_down_sync(0xffffffffu, a[q], o);
#pragma unroll
        for (int q = 0; q < NV; ++q) a[q] = red_combine(op[q], a[q], b[q]);
      }
      return;
    }
#pragma unroll
    for (int c0 = 0; c0 < NV; c0 += kChunk) {
#pragma unroll
      for (int q = 0; q < kChunk; ++q) {
        if (c0 + q < NV) tile[q][threadIdx.x] = a[c0 + q];
      }
      __syncwarp();
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        if (lane < o) {
#pragma unroll
          for (int q = 0; q < kChunk; ++q) {
            if (c0 + q < NV) {
              a[c0 + q] = red_combine(op[c0 + q], a[c0 + q], tile[q][threadIdx.x + o]);
              if (o > 1) tile[q][threadIdx.x] = a[c0 + q];
            }
          }
        }
        __syncwarp();
      }
    }
  };
  double a[NV];
#pragma unroll
  for (int q = 0; q < NV; ++q) a[q] = v[q];
  warp_tree(a);
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < NV; ++q) red[q][warp] = a[q];
  }
  __syncthreads();
  if (warp == 0) {
    const int nw = blockDim.x >> 5;
#pragma unroll
    for (int q = 0; q < NV; ++q) a[q] = lane < nw ? red[q][lane] : red_identity(op[q]);
    warp_tree(a);
    if (lane == 0) {
#pragma unroll
      for (int q = 0; q < NV; ++q) out[q] = a[q];
    }
  }
  __syncthreads();
}

/// Returns true in the (single) block that holds the grid-wide result in res.
/// Partials are stored quantity-major (partials[q·gridDim + block]): NV threads
/// of a block write them side by side and the last block reads them coalesced.
template <int NV>
__device__ __forceinline__ bool grid_reduce(double (&v)[NV], const int (&op)[NV],
                                            double* __restrict__ partials,
                                            unsigned int* __restrict__ counter, double* res) {
  __shared__ bool is_last;
  block_reduce<NV>(v, op, res);
  if (gridDim.x == 1) return true;
  if (threadIdx.x < NV) {
    partials[threadIdx.x * gridDim.x + blockIdx.x] = res[threadIdx.x];
    __threadfence();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // atomicInc wraps to 0 at gridDim.x − 1: the counter resets itself
    const unsigned int t = atomicInc(counter, gridDim.x - 1);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return false;
  __threadfence();
  double w[NV];
#pragma unroll
  for (int q = 0; q < NV; ++q) w[q] = red_identity(op[q]);
  for (int b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
    double p[NV];
#pragma unroll
    for (int q = 0; q < NV; ++q) p[q] = __ldcg(&partials[q * gridDim.x + b]);
#pragma unroll
    for (int q = 0; q < NV; ++q) w[q] = red_combine(op[q], w[q], p[q]);
  }
  block_reduce<NV>(w, op, res);
  return true;
}

struct RedBuf {
  double* partials;
  unsigned int* counter;
};

/// slpb_point_info for a point: vals = [f | c_e | c_i], slack s.
/// out: f, ce_l1, cis_l1, log_s_sum, finite bits (as double), ci_all_positive.
__global__ void __launch_bounds__(kReduceThreads)
k_point_info(const double* __restrict__ vals, const double* __restrict__ s,
             int me, int mi, RedBuf rb, double* __restrict__ out) {
  // 0 ce_l1, 1 cis_l1, 2 log sum, 3 nonfinite c_e, 4 nonfinite c_i, 5 c_i ≤ 0
  double w[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  const double* c_e = vals + 1;
  const double* c_i = vals + 1 + me;
  const int stride = gridDim.x * blockDim.x;
  const int t0 = blockIdx.x * blockDim.x + threadIdx.x;
  for (int i = t0; i < me; i += stride) {
    const double c = c_e[i];
    w[0] += fabs(c);
    if (!isfinite(c)) w[3] += 1.0;
  }
  for (int i = t0; i < mi; i += stride) {
    const double c = c_i[i];
    w[1] += fabs(c - s[i]);
    w[2] += log(s[i]);
    if (!isfinite(c)) w[4] += 1.0;
    if (!(c > 0.0)) w[5] += 1.0;
  }
  const int op[6] = {RED_SUM, RED_SUM, RED_SUM, RED_SUM, RED_SUM, RED_SUM};
  __shared__ double res[6];
  if (!grid_reduce<6>(w, op, rb.partials, rb.counter, res)) return;
  if (threadIdx.x == 0) {
    const double f = vals[0];
    int bits = 0;
    if (isfinite(f)) bits |= SLPB_FINITE_F;
    if (res[3] == 0.0) bits |= SLPB_FINITE_C_E;
    if (res[4] == 0.0) bits |= SLPB_FINITE_C_I;
    out[0] = f;
    out[1] = res[0];
    out[2] = res[1];
    out[3] = res[2];
    out[4] = static_cast<double>(bits);
    out[5] = res[5] == 0.0 ? 1.0 : 0.0;
  }
}

/// Finite-ness of the derivative arrays: out[0] = OR of G/A_E/A_I/H bits.
__global__ void __launch_bounds__(kReduceThreads)
k_deriv_finite(const double* __restrict__ dvals, int64_t off_ae, int64_t off_ai,
               int64_t off_h, int64_t total, RedBuf rb,
               double* __restrict__ out) {
  double w[4] = {0.0, 0.0, 0.0, 0.0};
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += stride) {
    if (!isfinite(dvals[i])) {
      if (i < off_ae) {
        w[0] += 1.0;
      } else if (i < off_ai) {
        w[1] += 1.0;
      } else if (i < off_h) {
        w[2] += 1.0;
      } else {
        w[3] += 1.0;
      }
    }
  }
  const int op[4] = {RED_SUM, RED_SUM, RED_SUM, RED_SUM};
  __shared__ double res[4];
  if (!grid_reduce<4>(w, op, rb.partials, rb.counter, res)) return;
  if (threadIdx.x == 0) {
    int bits = 0;
    if (res[0] == 0.0) bits |= SLPB_FINITE_G;
    if (res[1] == 0.0) bits |= SLPB_FINITE_A_E;
    if (res[2] == 0.0) bits |= SLPB_FINITE_A_I;
    if (res[3] == 0.0) bits |= SLPB_FINITE_H;
    out[0] = static_cast<double>(bits);
  }
}

struct CscView {
  const int32_t* colptr;
  const int32_t* rowidx;
  const double* val;
};

/// All reductions behind kkt_error / unscaled_kkt_error (kkt_error.hpp:92-251),
/// is_locally_infeasible.hpp and the divergence guard, for one point.
/// out layout = fields of slpb_kkt_stats in declaration order.
__global__ void __launch_bounds__(kReduceThreads)
k_kkt_stats(CscView Ae, CscView Ai, const double* __restrict__ g,
            const double* __restrict__ c_e, const double* __restrict__ c_i,
            const double* __restrict__ x, const double* __restrict__ s,
            const double* __restrict__ y, const double* __restrict__ z,
            const double* __restrict__ d_c,
            const double* __restrict__ inv_d_c, double d_f, double mu, int n,
            int me, int mi, RedBuf rb, double* __restrict__ out,
            const double* __restrict__ scan, int64_t off_ae, int64_t off_ai,
            int64_t off_h, int64_t scan_total, double* __restrict__ finite_out) {
  // (inv_d_c[r] = RN(1 / d_c[r]), computed once when the scaling is set: the
  // reference divides per entry, kkt_error.hpp:150-190 — same value, no
  // division inside the column loops)
  const double inv_d_f = 1.0 / d_f;
  // 0 r_inf 1 r_l1 2 y_l1 3 z_l1 4 sz_min 5 sz_max 6 sz_mu_l1 7 ce_inf 8 ce_l1
  // 9 cis_inf 10 cis_l1 11 u_r_inf 12 u_y_l1 13 u_z_l1 14 u_sz_min 15 u_sz_max
  // 16 u_ce_inf 17 u_cis_inf 18 aetce_sq 19 ce_sq 20 aitcip_sq 21 cip_sq
  // 22 x_inf 23 s_inf 24 nonfinite count
  // 25…28 (scan != nullptr: k_deriv_finite in the same launch) non-finite
  // entries of g, A_e, A_i, H
  double v[29];
#pragma unroll
  for (int q = 0; q < 29; ++q) v[q] = 0.0;
  v[4] = INFINITY;
  v[5] = -INFINITY;
  v[14] = INFINITY;
  v[15] = -INFINITY;
  const int stride = gridDim.x * blockDim.x;
  const int t0 = blockIdx.x * blockDim.x + threadIdx.x;
  for (int c = t0; c < n; c += stride) {
    // scaled: g − A_eᵀy − A_iᵀz ; unscaled: g/d_f − (A_e/d_ce)ᵀ(d_ce y/d_f) − …
    double aty = 0.0, atz = 0.0, u_aty = 0.0, u_atz = 0.0, atc = 0.0,
           atcp = 0.0;
    const int ae_end = Ae.colptr[c + 1], ai_end = Ai.colptr[c + 1];
    // (the gathers of four entries are in flight together; the sums keep the
    // entry order)
#pragma unroll 4
    for (int k = Ae.colptr[c]; k < ae_end; ++k) {
      const int r = Ae.rowidx[k];
      const double a = Ae.val[k];
      aty += a * y[r];
      const double dc = d_c[r];
      u_aty += (inv_d_c[r] * a) * (dc * y[r] * inv_d_f);
      atc += a * c_e[r];
    }
#pragma unroll 4
    for (int k = Ai.colptr[c]; k < ai_end; ++k) {
      const int r = Ai.rowidx[k];
      const double a = Ai.val[k];
      atz += a * z[r];
      const double dc = d_c[me + r];
      u_atz += (inv_d_c[me + r] * a) * (dc * z[r] * inv_d_f);
      atcp += a * fmin(c_i[r], 0.0);
    }
    const double r = g[c] - aty - atz;
    const double ur = inv_d_f * g[c] - u_aty - u_atz;
    v[0] = fmax(v[0], fabs(r));
    v[1] += fabs(r);
    v[11] = fmax(v[11], fabs(ur));
    v[18] += atc * atc;
    v[20] += atcp * atcp;
    const double xv = x[c];
    v[22] = fmax(v[22], fabs(xv));
    if (!isfinite(xv)) v[24] += 1.0;
  }
  for (int i = t0; i < me; i += stride) {
    const double c = c_e[i], yv = y[i], dc = d_c[i];
    v[2] += fabs(yv);
    v[7] = fmax(v[7], fabs(c));
    v[8] += fabs(c);
    v[12] += fabs(dc * yv * inv_d_f);
    v[16] = fmax(v[16], fabs(inv_d_c[i] * c));
    v[19] += c * c;
  }
  for (int i = t0; i < mi; i += stride) {
    const double c = c_i[i], sv = s[i], zv = z[i], dc = d_c[me + i];
    const double inv = inv_d_c[me + i];
    v[3] += fabs(zv);
    const double sz = sv * zv;
    v[4] = fmin(v[4], sz);
    v[5] = fmax(v[5], sz);
    v[6] += fabs(sz - mu);
    v[9] = fmax(v[9], fabs(c - sv));
    v[10] += fabs(c - sv);
    const double zu = dc * zv * inv_d_f;
    v[13] += fabs(zu);
    const double szu = (inv * sv) * zu;
    v[14] = fmin(v[14], szu);
    v[15] = fmax(v[15], szu);
    v[17] = fmax(v[17], fabs(inv * c - inv * sv));
    const double cp = fmin(c, 0.0);
    v[21] += cp * cp;
    v[23] = fmax(v[23], fabs(sv));
    if (!isfinite(sv)) v[24] += 1.0;
  }
  if (scan != nullptr) {
    const int64_t sstride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t i = t0; i < scan_total; i += sstride) {
      if (!isfinite(scan[i])) {
        // (static indices: v[] must stay in registers)
        if (i < off_ae) {
          v[25] += 1.0;
        } else if (i < off_ai) {
          v[26] += 1.0;
        } else if (i < off_h) {
          v[27] += 1.0;
        } else {
          v[28] += 1.0;
        }
      }
    }
  }
  const int op[29] = {RED_MAX, RED_SUM, RED_SUM, RED_SUM, RED_MIN, RED_MAX,
                      RED_SUM, RED_MAX, RED_SUM, RED_MAX, RED_SUM, RED_MAX,
                      RED_SUM, RED_SUM, RED_MIN, RED_MAX, RED_MAX, RED_MAX,
                      RED_SUM, RED_SUM, RED_SUM, RED_SUM, RED_MAX, RED_MAX,
                      RED_SUM, RED_SUM, RED_SUM, RED_SUM, RED_SUM};
  __shared__ double res[29];
  if (!grid_reduce<29>(v, op, rb.partials, rb.counter, res)) return;
  if (threadIdx.x == 0 && finite_out != nullptr) {
    int bits = 0;
    if (res[25] == 0.0) bits |= SLPB_FINITE_G;
    if (res[26] == 0.0) bits |= SLPB_FINITE_A_E;
    if (res[27] == 0.0) bits |= SLPB_FINITE_A_I;
    if (res[28] == 0.0) bits |= SLPB_FINITE_H;
    finite_out[0] = static_cast<double>(bits);
  }
  if (threadIdx.x == 0) {
    for (int q = 0; q < 18; ++q) out[q] = res[q];
    out[18] = sqrt(res[18]);
    out[19] = sqrt(res[19]);
    out[20] = sqrt(res[20]);
    out[21] = sqrt(res[21]);
    out[22] = res[22];
    out[23] = res[23];
    out[24] = res[24] == 0.0 ? 1.0 : 0.0;
  }
}

// ---------------------------------------------------------------------------
// kernels: KKT system
// ---------------------------------------------------------------------------

/// Σ = S⁻¹Z and the inequality part of the rhs (interior_point.hpp:426,447).
/// mode 0: t = −Σc_i + μS⁻¹e + z ; mode 1 (SOC, :615): t = μS⁻¹e − Σ·cis_soc
__global__ void k_sigma_t(const double* __restrict__ s,
                          const double* __restrict__ z,
                          const double* __restrict__ c_i,
                          const double* __restrict__ cis_soc, double mu,
                          int mode, int mi, double* __restrict__ sinv,
                          double* __restrict__ sigma,
                          double* __restrict__ t) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= mi) return;
  const double si = 1.0 / s[i];
  const double sg = si * z[i];
  sinv[i] = si;
  sigma[i] = sg;
  t[i] = mode == 0 ? (-sg * c_i[i] + mu * si + z[i])
                   : (mu * si - sg * cis_soc[i]);
}

/// rhs = −[g − A_eᵀy − A_iᵀt ; c_e] (interior_point.hpp:444-448) and k_sigma_t in
/// one launch: thread c < n + m_e forms its right-hand side entry with t recomputed entry by entry from the same expressions (so
/// nothing waits for the Σ/t arrays), thread i < m_i stores S⁻¹, Σ and t for
/// the kernels that follow (assembly, step recovery).
__global__ void k_rhs_sigma(CscView Ae, CscView Ai, const double* __restrict__ g,
                            const double* __restrict__ y,
                            const double* __restrict__ c_e,
                            const double* __restrict__ s,
                            const double* __restrict__ z,
                            const double* __restrict__ c_i,
                            const double* __restrict__ cis_soc, double mu,
                            int mode, int n, int me, int mi,
                            double* __restrict__ sinv, double* __restrict__ sigma,
                            double* __restrict__ t, double* __restrict__ rhs) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < mi) {
    const double si = 1.0 / s[c];
    const double sg = si * z[c];
    sinv[c] = si;
    sigma[c] = sg;
    t[c] = mode == 0 ? (-sg * c_i[c] + mu * si + z[c]) : (mu * si - sg * cis_soc[c]);
  }
  if (c < n) {
    double aty = 0.0, att = 0.0;
    for (int k = Ae.colptr[c]; k < Ae.colptr[c + 1]; ++k) {
      aty += Ae.val[k] * y[Ae.rowidx[k]];
    }
    for (int k = Ai.colptr[c]; k < Ai.colptr[c + 1]; ++k) {
      const int r = Ai.rowidx[k];
      const double si = 1.0 / s[r];
      const double sg = si * z[r];
      const double tr =
          mode == 0 ? (-sg * c_i[r] + mu * si + z[r]) : (mu * si - sg * cis_soc[r]);
      att += Ai.val[k] * tr;
    }
    rhs[c] = -g[c] + aty + att;
  } else if (c < n + me) {
    rhs[c] = -c_e[c - n];
  }
}

__global__ void k_kkt_assemble(const int32_t* __restrict__ h_idx,
                               const int32_t* __restrict__ ae_idx,
                               const int32_t* __restrict__ prod_ptr,
                               const int32_t* __restrict__ prod_a,
                               const int32_t* __restrict__ prod_b,
                               const int32_t* __restrict__ prod_row,
                               const double* __restrict__ Hv,
                               const double* __restrict__ Aev,
                               const double* __restrict__ Aiv,
                               const double* __restrict__ sigma, int nnz,
                               double* __restrict__ Kval) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nnz) return;
  Kval[e] = kkt_entry(e, h_idx, ae_idx, prod_ptr, prod_a, prod_b, prod_row, Hv,
                      Aev, Aiv, sigma);
}

/// k_kkt_assemble and k_init_factor in one launch (the first factorisation
/// attempt of an iteration needs both).
__global__ void k_kkt_assemble_init(const int32_t* __restrict__ h_idx,
                                    const int32_t* __restrict__ ae_idx,
                                    const int32_t* __restrict__ prod_ptr,
                                    const int32_t* __restrict__ prod_a,
                                    const int32_t* __restrict__ prod_b,
                                    const int32_t* __restrict__ prod_row,
                                    const double* __restrict__ Hv,
                                    const double* __restrict__ Aev,
                                    const double* __restrict__ Aiv,
                                    const double* __restrict__ sigma, int nnz,
                                    double* __restrict__ Kval,
                                    int32_t* __restrict__ stats, int32_t* sync,
                                    int ns, int zero_sync, int32_t* solve_sync) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (solve_sync != nullptr && e < 1 + 3 * ns) {
    solve_sync[e] = (e > ns && e <= 2 * ns) ? 1 : 0;  // see k_init_factor
  }
  if (e < 16) stats[e] = (e & 7) == 5 ? 0x7ff00000 : 0;
  if (zero_sync) {
    if (e < 1 + 2 * ns) sync[e] = 0;
    if (e < 2) sync[1 + 3 * ns + e] = 0;
  }
  if (e >= nnz) return;
  Kval[e] = kkt_entry(e, h_idx, ae_idx, prod_ptr, prod_a, prod_b, prod_row, Hv,
                      Aev, Aiv, sigma);
}

// ---- Lagrange multiplier estimate (lagrange_multiplier_estimate.hpp:55-131) ---
// The reference solves the normal equations (ÂÂᵀ)[y; z] = Â[∇f; −μe] with
// Â = [A_e 0; A_i −S]. Eliminating the slack block of the equivalent augmented
// system [I Âᵀ; Â 0] gives a system with the pattern of the Newton matrix:
//   [I + A_iᵀS⁻²A_i  A_eᵀ; A_e  0] [r; y] = [∇f − μA_iᵀS⁻¹e; 0],
//   z = S⁻²(A_i r + μs),
// so the estimate reuses the assembly recipe, the symbolic factorisation and the
// tree kernels (H → I, Σ → S⁻²).

/// lhs entry e with H replaced by the identity; sigma holds S⁻².
__global__ void k_kkt_assemble_identity(
    const uint8_t* __restrict__ diag_flag, const int32_t* __restrict__ ae_idx,
    const int32_t* __restrict__ prod_ptr, const int32_t* __restrict__ prod_a,
    const int32_t* __restrict__ prod_b, const int32_t* __restrict__ prod_row,
    const double* __restrict__ Aev, const double* __restrict__ Aiv,
    const double* __restrict__ sigma, int nnz, double* __restrict__ Kval) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nnz) return;
  const int ae = ae_idx[e];
  if (ae >= 0) {
    Kval[e] = Aev[ae];
    return;
  }
  double v = diag_flag[e] ? 1.0 : 0.0;
  for (int k = prod_ptr[e]; k < prod_ptr[e + 1]; ++k) {
    v += (Aiv[prod_a[k]] * sigma[prod_row[k]]) * Aiv[prod_b[k]];
  }
  Kval[e] = v;
}

__global__ void k_inv_square(const double* __restrict__ s, int mi,
                             double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < mi) out[i] = 1.0 / (s[i] * s[i]);
}

/// rhs = [∇f − μ A_iᵀ S⁻¹ e ; 0]
__global__ void k_rhs_estimate(CscView Ai, const double* __restrict__ g,
                               const double* __restrict__ s, double mu, int n,
                               int me, double* __restrict__ rhs) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n) {
    double acc = 0.0;
    for (int k = Ai.colptr[c]; k < Ai.colptr[c + 1]; ++k) {
      acc += Ai.val[k] * (1.0 / s[Ai.rowidx[k]]);
    }
    rhs[c] = g[c] - mu * acc;
  } else if (c < n + me) {
    rhs[c] = 0.0;
  }
}

/// y = sol₂ ; z = clamp(S⁻²(A_i r + μs), μ/(κs), κμ/s), κ = 1e10.
__global__ void k_estimate_finish(const double* __restrict__ sol,
                                  const int32_t* __restrict__ ai_rowptr,
                                  const int32_t* __restrict__ ai_rcol,
                                  const int32_t* __restrict__ ai_ridx,
                                  const double* __restrict__ ai_val,
                                  const double* __restrict__ s, double mu, int n,
                                  int me, int mi, double* __restrict__ y,
                                  double* __restrict__ z) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < me) y[i] = sol[n + i];
  if (i < mi) {
    double acc = 0.0;
    for (int k = ai_rowptr[i]; k < ai_rowptr[i + 1]; ++k) {
      acc += ai_val[ai_ridx[k]] * sol[ai_rcol[k]];
    }
    const double si = s[i];
    const double v = (acc + mu * si) / (si * si);
    const double kappa = 1e10;
    const double lo = 1.0 / kappa * mu / si;
    const double hi = kappa * mu / si;
    z[i] = v < lo ? lo : (hi < v ? hi : v);
  }
}

// ---------------------------------------------------------------------------
// kernels: multifrontal LDLᵀ, one thread block per front, one launch per
// level of the assembly tree
// ---------------------------------------------------------------------------

// (one block per front; the level kernels only see the rare fronts above order
// 32 and their ancestors, which are big: 16 warps)
constexpr int kFrontThreads = 512;

__global__ void k_factor_level(SymbolicView S,
                               const int32_t* __restrict__ level_supers,
                               const double* __restrict__ Kval, double delta,
                               double gamma, double* __restrict__ panels,
                               double* __restrict__ updates,
                               double* __restrict__ D,
                               int32_t* __restrict__ stats, int fused_arith,
                               double* gscratch, const int32_t* __restrict__ gslot,
                               int64_t gstride, int smem_cap_doubles) {
  extern __shared__ double smem[];
  __shared__ int ls[6];
  const int s = level_supers[blockIdx.x];
  const int F = S.front_dim[s];
  // a front beyond the shared-memory cap is eliminated in a workspace in
  // global memory (L2): slow, but any order works
  double* W = smem;
  if (F * F + F > smem_cap_doubles) W = gscratch + int64_t(gslot[s]) * gstride;
  double* lcol = W + size_t(F) * F;
  ldlt_factor_front<kFrontThreads>(threadIdx.x, s, S, Kval, delta, gamma,
                                   panels, updates, D, W, lcol, ls,
                                   BlockSync{}, fused_arith != 0);
  if (threadIdx.x == 0) {
    atomicAdd(&stats[0], ls[0]);
    atomicAdd(&stats[1], ls[1]);
    atomicAdd(&stats[2], ls[2]);
    atomicOr(&stats[3], ls[3]);
    const unsigned long long bits =
        (unsigned long long)(unsigned)ls[4] |
        ((unsigned long long)(unsigned)ls[5] << 32);
    atomicMin(reinterpret_cast<unsigned long long*>(&stats[4]), bits);
  }
}

__global__ void k_forward_level(SymbolicView S,
                                const int32_t* __restrict__ level_supers,
                                const double* __restrict__ panels,
                                const double* __restrict__ rhs,
                                double* __restrict__ xperm,
                                double* __restrict__ uvecs) {
  extern __shared__ double smem[];
  const int s = level_supers[blockIdx.x];
  ldlt_forward_front<kFrontThreads>(threadIdx.x, s, S, panels, rhs, xperm,
                                    uvecs, smem, BlockSync{});
}

__global__ void k_backward_level(SymbolicView S,
                                 const int32_t* __restrict__ level_supers,
                                 const double* __restrict__ panels,
                                 const double* __restrict__ D,
                                 double* __restrict__ xperm) {
  extern __shared__ double smem[];
  const int s = level_supers[blockIdx.x];
  ldlt_backward_front<kFrontThreads>(threadIdx.x, s, S, panels, D, xperm, smem,
                                     BlockSync{});
}

__global__ void k_unpermute(const double* __restrict__ xperm,
                            const int32_t* __restrict__ perm, int dim,
                            double* __restrict__ sol) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < dim) sol[perm[k]] = xperm[k];
}

// ---------------------------------------------------------------------------
// kernels: multifrontal LDLᵀ as ONE dependency-driven launch (fronts ≤ 32).
// One warp per front, front in shared memory, warp-level syncs only. Warps
// draw fronts from an atomic ticket counter in level order (children always
// hold smaller tickets, so whatever a warp waits for is already running or
// done: no deadlock, no grid-wide barrier) and signal their parent through a
// per-front counter in global memory.
// ---------------------------------------------------------------------------

// Warps per block of the tree kernels and resident blocks per SM. A warp keeps a
// 13.6 KB front workspace in shared memory and k_factor_tree needs 136
// registers per thread unconstrained; four blocks of four warps per SM (128
// registers, 4 bytes of spill, 218 KB of shared memory) measured best on B200:
// 0.0953 ms per factor launch at N=5000 against 0.0978 (3 × 4 warps, 136
// registers), 0.0977 (6 × 2) and 0.1054 (blocks of 8).
#ifndef SLPB_TREE_WARPS
#define SLPB_TREE_WARPS 4
#endif
constexpr int kTreeWarps = SLPB_TREE_WARPS;

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

/// Spins until *p ≥ need, backing off so that idle warps do not flood the L2.
__device__ __forceinline__ void wait_at_least(const int* p, int need) {
  unsigned ns = 64;
  while (ld_acquire_gpu(p) < need) {
    __nanosleep(ns);
    if (ns < 1024) ns <<= 1;
  }
}

/// Publishes everything this warp wrote (the lanes were joined by __syncwarp
/// before) and bumps a dependency counter: one release-RMW instead of a full
/// fence followed by an atomic.
__device__ __forceinline__ void red_release_add(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v)
               : "memory");
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v)
               : "memory");
}

struct TreeView {
  const int32_t* order;      // fronts to process, by ascending level
  int32_t n_order;           // how many (single GPU: all n_super of them)
  const int32_t* order_bwd;  // k_solve_tree, backward pass: walked in reverse
  int32_t n_fwd, n_bwd;      // fronts of the forward / backward pass
  const FrontMeta* metas;    // one packed record per front
  const int32_t* child_idx;
  const int32_t* rel_idx;
  const int32_t* rows_idx;
  const int32_t* asm_src;
  const int32_t* asm_dst;    // position in the warp's front (leading dim kFrontLd)
  ExtendList ext;            // per-front extend-add lists (ldlt_warp.cuh)
  const uint8_t* col_is_primal;
  const int32_t* perm;
  int32_t* sync;             // [0] ticket | fcount[ns] | fflag[ns] | bflag[ns]
  int32_t n_super;
  unsigned long long* debug; // optional: 4 globaltimer stamps per front
};

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

/// Second regularisation of a speculative pair (slpb_factor_pair) and where its
/// results go: variant v of the factorisation writes panels + v·panel_stride,
/// updates + v·update_stride, D + v·dim and stats + 8·v.
struct FactorPair {
  int n_variants;
  double delta1, gamma1;
  int64_t panel_stride, update_stride, uvec_stride;
  int32_t dim;
  const double* rhs;  // non-null: carry the forward substitution of this rhs
  double* xperm;      // variant v: xperm + v·dim
  double* uvecs;      // variant v: uvecs + v·uvec_stride
  int fused_arith;    // SLPB_ARITH_TENSOR: fused Schur updates, DMMA on dense fronts
};

#ifndef SLPB_TREE_MIN_BLOCKS
#define SLPB_TREE_MIN_BLOCKS 4
#endif
__global__ void __launch_bounds__(kTreeWarps * 32, SLPB_TREE_MIN_BLOCKS)
k_factor_tree(TreeView T, const double* __restrict__ Kval, double delta,
              double gamma, FactorPair pair, double* __restrict__ panels,
              double* updates, double* __restrict__ D,
              int32_t* __restrict__ stats) {
  extern __shared__ double smem[];
  __shared__ int ls[kTreeWarps][6];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* W = smem + size_t(warp) * kFrontSmemDoubles;
  const int nv = pair.n_variants;
  int acc_pos[2] = {0, 0}, acc_neg[2] = {0, 0}, acc_zero[2] = {0, 0};
  int acc_zpiv[2] = {0, 0};
  unsigned long long acc_min[2] = {0x7ff0000000000000ull,
                                   0x7ff0000000000000ull};  // +inf
  for (;;) {
    int t = 0;
    if (lane == 0) t = atomicAdd(&T.sync[0], 1);
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t >= nv * T.n_order) break;
    // both variants of a front hold neighbouring tickets: children (of either
    // variant) always hold smaller tickets than their parents
    const int v = nv == 2 ? (t & 1) : 0;
    const int s = T.order[nv == 2 ? (t >> 1) : t];
    const FrontMeta fm = load_front_meta(T.metas + s);
    int32_t* fcount = T.sync + 1 + v * T.n_super;
    int32_t* vstats = stats + 8 * v;
    // A variant that met an exactly zero pivot has failed whatever follows
    // (SimplicialLDLT reports NumericalIssue; the host only looks at the flag):
    // its remaining fronts just hand the dependency on. For the speculated
    // pair this makes the doomed (0, 0) attempt nearly free.
    int32_t* dead = T.sync + 1 + 3 * T.n_super + v;
    if (ld_acquire_gpu(dead) != 0) {
      if (lane == 0 && fm.parent >= 0) red_release_add(&fcount[fm.parent], 1);
      continue;
    }
    if (T.debug && lane == 0 && v == 0) T.debug[4 * s] = global_ns();
    const FusedRhs fr{pair.rhs, T.perm, pair.xperm + size_t(v) * pair.dim,
                      pair.uvecs + v * pair.uvec_stride};
    ldlt_factor_front_warp(
        lane, fm, T.asm_src, T.asm_dst, T.ext, T.col_is_primal, Kval,
        v ? pair.delta1 : delta,
        v ? pair.gamma1 : gamma, panels + v * pair.panel_stride,
        updates + v * pair.update_stride, D + v * pair.dim, W,
        &fcount[s], ls[warp], fr, pair.fused_arith != 0,
        (T.debug && v == 0) ? T.debug + 4 * s + 1 : nullptr);
    __syncwarp();
    if (T.debug && lane == 0 && v == 0) T.debug[4 * s + 3] = global_ns();
    if (lane == 0) {
      // hand over to the parent first; the inertia bookkeeping stays in
      // registers until the warp runs out of fronts
      if (fm.parent >= 0) red_release_add(&fcount[fm.parent], 1);
      acc_pos[v] += ls[warp][0];
      acc_neg[v] += ls[warp][1];
      acc_zero[v] += ls[warp][2];
      acc_zpiv[v] |= ls[warp][3];
      if (ls[warp][3]) {
        atomicExch(dead, 1);
        atomicOr(&vstats[3], 1);  // reported even if this warp is preempted
      }
      const unsigned long long bits =
          (unsigned long long)(unsigned)ls[warp][4] |
          ((unsigned long long)(unsigned)ls[warp][5] << 32);
      acc_min[v] = bits < acc_min[v] ? bits : acc_min[v];
    }
    (void)vstats;
  }
  if (lane == 0) {
    for (int v = 0; v < nv; ++v) {
      int32_t* vstats = stats + 8 * v;
      if (acc_pos[v]) atomicAdd(&vstats[0], acc_pos[v]);
      if (acc_neg[v]) atomicAdd(&vstats[1], acc_neg[v]);
      if (acc_zero[v]) atomicAdd(&vstats[2], acc_zero[v]);
      if (acc_zpiv[v]) atomicOr(&vstats[3], acc_zpiv[v]);
      atomicMin(reinterpret_cast<unsigned long long*>(&vstats[4]), acc_min[v]);
    }
  }
}

/// Prepares a factor launch in one go: the inertia statistics of both variants
/// (zero counts, min|D| = +inf) and, for the tree kernel, the ticket, the
/// dependency counters of both variants and the "variant is dead" flags.
__global__ void k_init_factor(int32_t* __restrict__ stats, int32_t* sync,
                              int ns, int zero_sync, int32_t* solve_sync) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  // (solve_sync != nullptr: the words of the solve that will follow, preset
  // for "forward substitution already carried by the factorisation" — what
  // k_init_solve_sync(…, skip_forward = 1) writes)
  if (solve_sync != nullptr && i < 1 + 3 * ns) {
    solve_sync[i] = (i > ns && i <= 2 * ns) ? 1 : 0;
  }
  if (i < 16) {
    // per variant: n_pos n_neg n_zero zero_pivot | +inf bits | pad
    const int k = i & 7;
    stats[i] = k == 5 ? 0x7ff00000 : 0;
  }
  if (zero_sync) {
    if (i < 1 + 2 * ns) sync[i] = 0;
    if (i < 2) sync[1 + 3 * ns + i] = 0;
  }
}

/// Prepares the ticket/flag words of k_solve_tree. With skip_forward the
/// forward substitution was carried by the factorisation: every "forward
/// done" flag is already set. fcount_init (optional): children of each front
/// that are complete before the launch (sharded solves: the subtree roots that
/// arrived through the exchange).
__global__ void k_init_solve_sync(int32_t* __restrict__ sync, int ns,
                                  int skip_forward,
                                  const int32_t* __restrict__ fcount_init,
                                  const int32_t* __restrict__ bflag_init = nullptr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) sync[0] = 0;
  if (i < ns) {
    sync[1 + i] = fcount_init ? fcount_init[i] : 0;  // fcount
    sync[1 + ns + i] = skip_forward ? 1 : 0;         // fflag
    // bflag (hybrid: the fronts of the top, solved by the level kernels, count
    // as done)
    sync[1 + 2 * ns + i] = bflag_init ? bflag_init[i] : 0;
  }
}

/// Forward (leaves→roots) then backward (roots→leaves) substitution in one
/// launch; tickets [0, n_fwd) are forward fronts (T.order), the next n_bwd
/// backward fronts (T.order_bwd in reverse). The backward pass writes the
/// un-permuted solution directly.
__global__ void __launch_bounds__(kTreeWarps * 32)
k_solve_tree(TreeView T, const double* __restrict__ panels,
             const double* __restrict__ D, const double* __restrict__ rhs,
             double* xperm, double* uvecs, double* __restrict__ sol) {
  __shared__ double wbuf[kTreeWarps][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* w = wbuf[warp];
  const int ns = T.n_super;
  int32_t* fcount = T.sync + 1;
  int32_t* fflag = fcount + ns;
  int32_t* bflag = fflag + ns;
  for (;;) {
    int t = 0;
    if (lane == 0) t = atomicAdd(&T.sync[0], 1);
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t >= T.n_fwd + T.n_bwd) break;
    if (t < T.n_fwd) {
      const int s = T.order[t];
      const FrontMeta fm = load_front_meta(T.metas + s);
      ldlt_forward_front_warp(lane, fm, T.metas, T.child_idx, T.rel_idx, T.perm,
                              panels, rhs, xperm, uvecs, w, &fcount[s]);
      __syncwarp();
      if (lane == 0) {
        if (fm.parent >= 0) {
          red_release_add(&fcount[fm.parent], 1);
        }
        st_release(&fflag[s], 1);
      }
    } else {
      const int s = T.order_bwd[T.n_fwd + T.n_bwd - 1 - t];
      const FrontMeta fm = load_front_meta(T.metas + s);
      ldlt_backward_front_warp(lane, fm, T.rows_idx, panels, D, xperm,
                               fm.parent >= 0 ? &bflag[fm.parent] : &fflag[s]);
      __syncwarp();
      if (lane < fm.np) sol[T.perm[fm.c0 + lane]] = xperm[fm.c0 + lane];
      __syncwarp();
      if (lane == 0) st_release(&bflag[s], 1);
    }
  }
}

// ---- multi-GPU sharded factorisation: exchange of the subtree roots ----------
// Segment of one rank in the exchange buffer, per variant:
//   [8 doubles of inertia statistics | len doubles of root data]
// pos[i]: bit 60 set → update-vector entry, else update-matrix entry.
constexpr int64_t kPackVecTag = int64_t(1) << 60;
constexpr int kPackStats = 8;

__global__ void k_tree_pack(const int64_t* __restrict__ pos, int len, int seg,
                            int nv, const double* __restrict__ updates,
                            int64_t update_stride,
                            const double* __restrict__ uvecs,
                            int64_t uvec_stride,
                            const int32_t* __restrict__ stats,
                            double* __restrict__ buf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int v = blockIdx.y;
  if (v >= nv) return;
  double* out = buf + size_t(v) * seg;
  if (i < kPackStats) {
    // n_pos n_neg n_zero zero_pivot as numbers, min|D| as it is
    double x = 0.0;
    if (i < 4) {
      x = static_cast<double>(stats[8 * v + i]);
    } else if (i == 4) {
      x = __longlong_as_double(
          *reinterpret_cast<const long long*>(&stats[8 * v + 4]));
    }
    out[i] = x;
  }
  if (i < len) {
    const int64_t p = pos[i];
    out[kPackStats + i] = (p & kPackVecTag)
                              ? uvecs[v * uvec_stride + (p & ~kPackVecTag)]
                              : updates[v * update_stride + p];
  }
}

/// After the all-gather: the other ranks' root data goes where the top fronts
/// expect it, their inertia counts are added to this rank's.
__global__ void k_tree_unpack(const int64_t* __restrict__ pos_all,
                              const int32_t* __restrict__ len_all, int max_len,
                              int seg, int nv, int world, int rank,
                              const double* __restrict__ buf,
                              double* __restrict__ updates,
                              int64_t update_stride, double* __restrict__ uvecs,
                              int64_t uvec_stride, int32_t* __restrict__ stats,
                              int with_stats) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y / nv, v = blockIdx.y % nv;
  if (r == rank || r >= world) return;
  const double* in = buf + (size_t(r) * nv + v) * seg;
  if (with_stats && i == 0) {
    atomicAdd(&stats[8 * v + 0], static_cast<int>(in[0]));
    atomicAdd(&stats[8 * v + 1], static_cast<int>(in[1]));
    atomicAdd(&stats[8 * v + 2], static_cast<int>(in[2]));
    atomicOr(&stats[8 * v + 3], static_cast<int>(in[3]));
    atomicMin(reinterpret_cast<unsigned long long*>(&stats[8 * v + 4]),
              static_cast<unsigned long long>(__double_as_longlong(in[4])));
  }
  if (i < len_all[r]) {
    const int64_t p = pos_all[size_t(r) * max_len + i];
    const double x = in[kPackStats + i];
    if (p & kPackVecTag) {
      uvecs[v * uvec_stride + (p & ~kPackVecTag)] = x;
    } else {
      updates[v * update_stride + p] = x;
    }
  }
}

__global__ void k_gather_idx(const double* __restrict__ src,
                             const int32_t* __restrict__ idx, int len,
                             double* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < len) dst[i] = src[idx[i]];
}
/// After the all-gather: sol[idx[r][i]] = buf[r][i] for every other rank.
__global__ void k_scatter_idx(const double* __restrict__ buf,
                              const int32_t* __restrict__ idx_all,
                              const int32_t* __restrict__ len_all, int max_len,
                              int world, int rank, double* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (r == rank || r >= world || i >= len_all[r]) return;
  dst[idx_all[size_t(r) * max_len + i]] = buf[size_t(r) * max_len + i];
}

// ---------------------------------------------------------------------------
// kernels: step recovery, line search
// ---------------------------------------------------------------------------

/// Step recovery (interior_point.hpp:470-481: pˢ = (c_i − s) + A_i pˣ,
/// pᶻ = μ/s − z − Σpˢ) and, in the same launch, the fraction-to-the-boundary
/// rule on (s, pˢ) and (z, pᶻ) (fraction_to_the_boundary_rule.hpp:19-43:
/// α = min(1, min −τ/pᵢ·xᵢ over the blocking components)), gᵀpˣ, (S⁻¹e)ᵀpˢ and
/// the step norms: every thread recovers the entries of its grid-stride slice
/// and feeds them straight into the reductions.
__global__ void __launch_bounds__(kReduceThreads)
k_step_recover_stats(const double* __restrict__ sol,
                     const int32_t* __restrict__ ai_rowptr,
                     const int32_t* __restrict__ ai_rcol,
                     const int32_t* __restrict__ ai_ridx,
                     const double* __restrict__ ai_val,
                     const double* __restrict__ c_i,
                     const double* __restrict__ s, const double* __restrict__ z,
                     const double* __restrict__ cis_soc,
                     const double* __restrict__ sinv,
                     const double* __restrict__ sigma,
                     const double* __restrict__ g, double mu, double tau, int n,
                     int me, int mi, double* __restrict__ px,
                     double* __restrict__ py, double* __restrict__ ps,
                     double* __restrict__ pz, RedBuf rb,
                     double* __restrict__ out) {
  // 0 alpha_max 1 alpha_z 2 g·px 3 sinv·ps 4..7 inf norms 8 nonfinite
  double v[9] = {1.0, 1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  const int stride = gridDim.x * blockDim.x;
  const int t0 = blockIdx.x * blockDim.x + threadIdx.x;
  for (int i = t0; i < n; i += stride) {
    const double p = sol[i];
    px[i] = p;
    v[2] += g[i] * p;
    v[4] = fmax(v[4], fabs(p));
    if (!isfinite(p)) v[8] += 1.0;
  }
  for (int i = t0; i < me; i += stride) {
    const double p = -sol[n + i];
    py[i] = p;
    v[6] = fmax(v[6], fabs(p));
  }
  for (int i = t0; i < mi; i += stride) {
    double acc = 0.0;
    for (int k = ai_rowptr[i]; k < ai_rowptr[i + 1]; ++k) {
      acc += ai_val[ai_ridx[k]] * sol[ai_rcol[k]];
    }
    const double cis = cis_soc ? cis_soc[i] : (c_i[i] - s[i]);
    const double p = cis + acc;
    const double q = mu * sinv[i] - z[i] - sigma[i] * p;
    ps[i] = p;
    pz[i] = q;
    if (p < 0.0) {
      const double cand = -tau / p * s[i];
      if (cand < v[0]) v[0] = cand;
    }
    if (q < 0.0) {
      const double cand = -tau / q * z[i];
      if (cand < v[1]) v[1] = cand;
    }
    v[3] += sinv[i] * p;
    v[5] = fmax(v[5], fabs(p));
    v[7] = fmax(v[7], fabs(q));
    if (!isfinite(p) || !isfinite(q)) v[8] += 1.0;
  }
  const int op[9] = {RED_MIN, RED_MIN, RED_SUM, RED_SUM, RED_MAX,
                     RED_MAX, RED_MAX, RED_MAX, RED_SUM};
  __shared__ double res[9];
  if (!grid_reduce<9>(v, op, rb.partials, rb.counter, res)) return;
  if (threadIdx.x == 0) {
    for (int q = 0; q < 8; ++q) out[q] = res[q];
    out[8] = res[8] == 0.0 ? 1.0 : 0.0;
  }
}

/// trial = iterate + α (p_x, p_s) and + α_z (p_y, p_z); writes trial x into
/// the trial leaf vector as well.
__global__ void k_trial_point(const double* __restrict__ x,
                              const double* __restrict__ s,
                              const double* __restrict__ y,
                              const double* __restrict__ z,
                              const double* __restrict__ px,
                              const double* __restrict__ ps,
                              const double* __restrict__ py,
                              const double* __restrict__ pz, double alpha,
                              double alpha_z,
                              const double* __restrict__ alpha_dev,
                              int dual_uses_primal, int n, int me, int mi,
                              double* __restrict__ tx, double* __restrict__ ts,
                              double* __restrict__ ty, double* __restrict__ tz,
                              double* __restrict__ leaf_trial) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (alpha_dev != nullptr) {
    // α_max, α_z as k_step_recover_stats just left them (slpb_solve_trial)
    alpha = alpha_dev[0];
    alpha_z = dual_uses_primal ? alpha_dev[0] : alpha_dev[1];
  }
  if (i < n) {
    const double v = x[i] + alpha * px[i];
    tx[i] = v;
    leaf_trial[i] = v;
  }
  if (i < mi) {
    ts[i] = s[i] + alpha * ps[i];
    tz[i] = z[i] + alpha_z * pz[i];
  }
  if (i < me) ty[i] = y[i] + alpha_z * py[i];
}

__global__ void k_copy(const double* __restrict__ src, double* __restrict__ dst,
                       int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}

/// Commit: z ← clamp(z, μ/(κ s), κ μ / s), κ = 1e10 (interior_point.hpp:797-801).
__global__ void k_clamp_z(const double* __restrict__ s, double mu, int mi,
                          double* __restrict__ z) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= mi) return;
  const double kappa = 1e10;
  const double lo = 1.0 / kappa * mu / s[i];
  const double hi = kappa * mu / s[i];
  const double v = z[i];
  z[i] = v < lo ? lo : (hi < v ? hi : v);  // std::clamp
}

/// The whole commit in one launch (interior_point.hpp:779-805): x, s, y, z and
/// the values f, c_e, c_i ← trial; z clamped like k_clamp_z (same expressions);
/// and the leaves of the accepted point [x ‖ d_ce⊙y ‖ d_ci⊙z] for the
/// re-linearisation that follows (k_prepare_leaves). Replaces five device
/// copies and two kernels.
__global__ void k_accept(const double* __restrict__ tx,
                         const double* __restrict__ ts,
                         const double* __restrict__ ty,
                         const double* __restrict__ tz,
                         const double* __restrict__ vals_trial,
                         const double* __restrict__ d_c, double mu, int n, int me,
                         int mi, double* __restrict__ x, double* __restrict__ s,
                         double* __restrict__ y, double* __restrict__ z,
                         double* __restrict__ vals_cur,
                         double* __restrict__ leaf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const double v = tx[i];
    x[i] = v;
    leaf[i] = v;
  }
  if (i < me) {
    const double v = ty[i];
    y[i] = v;
    leaf[n + i] = d_c[i] * v;
  }
  if (i < mi) {
    const double si = ts[i];
    s[i] = si;
    const double kappa = 1e10;
    const double lo = 1.0 / kappa * mu / si;
    const double hi = kappa * mu / si;
    const double v = tz[i];
    const double zc = v < lo ? lo : (hi < v ? hi : v);  // std::clamp
    z[i] = zc;
    leaf[n + me + i] = d_c[me + i] * zc;
  }
  if (i < 1 + me + mi) vals_cur[i] = vals_trial[i];
}

__global__ void k_soc_begin(const double* __restrict__ c_e,
                            const double* __restrict__ c_i,
                            const double* __restrict__ s, int me, int mi,
                            double* __restrict__ ce_soc,
                            double* __restrict__ cis_soc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < me) ce_soc[i] = c_e[i];
  if (i < mi) cis_soc[i] = c_i[i] - s[i];
}

/// c_e^soc ← α c_e^soc + trial c_e ; cis^soc ← α cis^soc + trial c_i − trial s
/// (interior_point.hpp:611-612)
__global__ void k_soc_accumulate(const double* __restrict__ tce,
                                 const double* __restrict__ tci,
                                 const double* __restrict__ ts, double alpha,
                                 int me, int mi, double* __restrict__ ce_soc,
                                 double* __restrict__ cis_soc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < me) ce_soc[i] = alpha * ce_soc[i] + tce[i];
  if (i < mi) cis_soc[i] = alpha * cis_soc[i] + tci[i] - ts[i];
}

// ---------------------------------------------------------------------------
// host-side launch helpers
// ---------------------------------------------------------------------------

inline int blocks_for(int64_t n, int threads) {
  return static_cast<int>((n + threads - 1) / threads);
}

/// Grid of a reduction kernel: about one element per thread (the reductions
/// are latency-bound: short per-thread loops matter more than few blocks), at
/// most two blocks per SM.
inline int red_blocks(int64_t n) {
  return std::max(1, std::min<int>(kReduceBlocks,
                                   blocks_for(n, kReduceThreads)));
}
inline RedBuf red_buf(slpb_solver* S) {
  return {S->red_partials.p, S->red_counter.p};
}

/// Shared memory one task may use: two tasks of the largest class stay
/// co-resident on an SM (228 KB per SM, 1 KB reserved per block).
constexpr int kAdSmemBudget = 112 * 1024;
constexpr int kAdSmemMax = 224 * 1024;

int upload_program_set(slpb_solver* S, ProgramSet& ps, DevProgramSet& d) {
  int budget = kAdSmemBudget;
  if (ps.max_smem > budget) budget = kAdSmemMax;  // one cluster per block
  if (const char* e = std::getenv("SLPB_AD_SMEM_BUDGET")) budget = std::atoi(e);
  if (!build_task_plan(ps, budget, S->error)) return SLPB_ERR_UNSUPPORTED;
  d.n_tasks = static_cast<int>(ps.task_prog.size());
  d.launches = ps.launches;
  d.max_smem = 0;
  for (const auto& L : d.launches) d.max_smem = std::max(d.max_smem, L.smem_bytes);
  CU(d.blob.upload(ps.blob, S->stream));
  CU(d.prog_offset.upload(ps.prog_offset, S->stream));
  CU(d.task_prog.upload(ps.task_prog, S->stream));
  CU(d.task_count.upload(ps.task_count, S->stream));
  CU(d.task_lanes.upload(ps.task_lanes, S->stream));
  CU(d.task_bind.upload(ps.task_bind, S->stream));
  CU(d.task_bindings.upload(ps.task_bindings, S->stream));
  return SLPB_OK;
}

/// Raises (never lowers) a kernel's dynamic shared-memory limit.
template <typename Kernel>
cudaError_t raise_dynamic_smem(Kernel kernel, int bytes) {
  static std::mutex mu;
  static std::map<const void*, int> current;
  std::lock_guard<std::mutex> lock{mu};
  const void* key = reinterpret_cast<const void*>(kernel);
  int& have = current[key];
  const int want = std::max(bytes, 48 * 1024);
  if (want <= have) return cudaSuccess;
  const cudaError_t e = cudaFuncSetAttribute(
      kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, want);
  if (e == cudaSuccess) have = want;
  return e;
}

/// Pinned result mirrors are recycled across handles: cudaFreeHost
/// synchronises the whole context and was seen to take 250 ms at the end of a
/// solve; a handle needs 512 bytes, so the few buffers a process ever creates
/// are simply kept.
struct PinnedPool {
  std::mutex mu;
  std::vector<double*> free_list;
  double* acquire() {
    {
      std::lock_guard<std::mutex> lock{mu};
      if (!free_list.empty()) {
        double* p = free_list.back();
        free_list.pop_back();
        return p;
      }
    }
    double* p = nullptr;
    if (cudaMallocHost(&p, kResultDoubles * sizeof(double)) != cudaSuccess) {
      return nullptr;
    }
    return p;
  }
  void give_back(double* p) {
    std::lock_guard<std::mutex> lock{mu};
    free_list.push_back(p);
  }
};
inline PinnedPool& pinned_pool() {
  static PinnedPool* pool = new PinnedPool;  // intentionally never destroyed
  return *pool;
}

int upload_gather(slpb_solver* S, const Gather& g, DevGather& d) {
  d.n_entries = g.n_entries();
  CU(d.ptr.upload(g.ptr, S->stream));
  CU(d.src_idx.upload(g.src_idx, S->stream));
  CU(d.src_scale.upload(g.src_scale, S->stream));
  std::vector<int32_t> longs;
  for (int e = 0; e < d.n_entries; ++e) {
    if (g.ptr[e + 1] - g.ptr[e] > kLongGather) longs.push_back(e);
  }
  d.n_long = static_cast<int>(longs.size());
  CU(d.long_entries.upload(longs, S->stream));
  return SLPB_OK;
}

int run_sweep(slpb_solver* S, const DevProgramSet& d, const double* leaf,
              double* stage) {
  if (d.n_tasks == 0) return SLPB_OK;
  const AdTasks A{d.blob.p,      d.prog_offset.p, d.task_prog.p,
                  d.task_count.p, d.task_lanes.p, d.task_bind.p,
                  d.task_bindings.p};
  // The launches of a sweep (one per shape of program class) are independent:
  // the first stays on the solver's stream, the others go to side streams
  // between a fork and a join, so that the small grids of the boundary rows
  // overlap the big one instead of queueing behind it.
  const size_t n_launch = d.launches.size();
  const bool fork = n_launch > 1 && !S->serial_sweeps;
  if (fork) CU(cudaEventRecord(S->fork_ev, S->stream));
  for (size_t i = 0; i < n_launch; ++i) {
    const auto& L = d.launches[i];
    cudaStream_t st = S->stream;
    if (fork && i > 0) {
      st = S->side[(i - 1) % slpb_solver::kSideStreams];
      CU(cudaStreamWaitEvent(st, S->fork_ev, 0));
    }
    k_ad_sweep<<<L.n_tasks, L.threads, L.smem_bytes, st>>>(A, L.first_task,
                                                            leaf, stage);
    ++S->counters.kernel_launches;
  }
  if (fork) {
    const size_t used = std::min<size_t>(n_launch - 1, slpb_solver::kSideStreams);
    for (size_t i = 0; i < used; ++i) {
      CU(cudaEventRecord(S->join_ev[i], S->side[i]));
      CU(cudaStreamWaitEvent(S->stream, S->join_ev[i], 0));
    }
  }
  CU(cudaGetLastError());
  return SLPB_OK;
}

/// The derivative sweep on `world` GPUs: this rank evaluates its share of the
/// tasks, then ONE NCCL all-gather over NVLink exchanges the stage slots every
/// rank produced (a few MB per Newton iteration), issued on the solver's stream
/// right behind the producing kernel.
int run_sweep_sharded(slpb_solver* S, const DevProgramSet& d, const double* leaf,
                      double* stage) {
  const ShardPlan& plan = S->shard;
  const int W = plan.world, r = S->rank;
  const AdTasks A{d.blob.p,      d.prog_offset.p, d.task_prog.p,
                  d.task_count.p, d.task_lanes.p, d.task_bind.p,
                  d.task_bindings.p};
  for (size_t L = 0; L < d.launches.size(); ++L) {
    const int cnt = plan.n_tasks[L * W + r];
    if (cnt == 0) continue;
    k_ad_sweep<<<cnt, d.launches[L].threads, d.launches[L].smem_bytes,
                 S->stream>>>(A, plan.first_task[L * W + r], leaf, stage);
    ++S->counters.kernel_launches;
  }
  CU(cudaGetLastError());
  if (plan.max_len == 0) return SLPB_OK;
  const bool timed = !S->cpending[0];
  if (timed) CU(cudaEventRecord(S->cev[0], S->stream));
  double* mine = S->shard_buf.p + size_t(r) * plan.max_len;
  if (plan.len[r] > 0) {
    k_shard_pack<<<blocks_for(plan.len[r], 256), 256, 0, S->stream>>>(
        stage, S->shard_slots.p + size_t(r) * plan.max_len, plan.len[r], mine);
    ++S->counters.kernel_launches;
  }
  const int rc = nccl_api().AllGather(mine, S->shard_buf.p, plan.max_len,
                                      kNcclFloat64, S->comm, S->stream);
  if (rc != 0) {
    return fail(S, SLPB_ERR_NCCL,
                std::string("ncclAllGather: ") +
                    (nccl_api().GetErrorString ? nccl_api().GetErrorString(rc)
                                               : "error"));
  }
  k_shard_unpack<<<blocks_for(int64_t(W) * plan.max_len, 256), 256, 0,
                   S->stream>>>(S->shard_buf.p, S->shard_slots.p, plan.max_len,
                                W, r, stage);
  if (timed) {
    CU(cudaEventRecord(S->cev[1], S->stream));
    S->cpending[0] = true;
    ++S->comm_timed[0];
  }
  ++S->comm_stats.count[0];
  S->comm_stats.bytes[0] += int64_t(plan.max_len) * 8;
  ++S->counters.kernel_launches;
  CU(cudaGetLastError());
  return SLPB_OK;
}

int run_gather(slpb_solver* S, const DevGather& d, const double* stage,
               double* out) {
  if (d.n_entries == 0) return SLPB_OK;
  k_gather<<<blocks_for(d.n_entries, 256), 256, 0, S->stream>>>(
      d.ptr.p, d.src_idx.p, d.src_scale.p, stage, S->d_f, S->d_c.p, out,
      d.n_entries);
  ++S->counters.kernel_launches;
  if (d.n_long > 0) {
    k_gather_long<<<d.n_long, 1024, 0, S->stream>>>(
        d.long_entries.p, d.ptr.p, d.src_idx.p, d.src_scale.p, stage, S->d_f,
        S->d_c.p, out);
    ++S->counters.kernel_launches;
  }
  CU(cudaGetLastError());
  return SLPB_OK;
}

/// After a stream synchronisation: folds the event pairs that were recorded
/// since the last harvest into the per-phase timers.
/// Start / end of timed kernel group w (0 eval_full, 1 eval_values, 2 assemble,
/// 3 factor, 4 solve): events are recorded for one launch in timer_every, and
/// never while the previous sample of the group has not been read yet.
cudaError_t timer_begin(slpb_solver* S, int w) {
  ++S->timers.launches[w];
  S->timing_now[w] = S->timer_every > 0 && !S->pending[w] &&
                     (S->timer_tick[w]++ % unsigned(S->timer_every)) == 0;
  return S->timing_now[w] ? cudaEventRecord(S->ev[2 * w], S->stream) : cudaSuccess;
}
cudaError_t timer_end(slpb_solver* S, int w) {
  if (!S->timing_now[w]) return cudaSuccess;
  S->timing_now[w] = false;
  S->pending[w] = true;
  return cudaEventRecord(S->ev[2 * w + 1], S->stream);
}

void harvest_timers(slpb_solver* S) {
  harvest_comm_timers(S);
  for (int w = 0; w < 5; ++w) {
    if (!S->pending[w]) continue;
    float ms = 0.0f;
    if (cudaEventElapsedTime(&ms, S->ev[2 * w], S->ev[2 * w + 1]) ==
        cudaSuccess) {
      S->last_ms[w] = ms;
      S->timers.total_ms[w] += ms;
      ++S->timers.count[w];
    } else {
      cudaGetLastError();
    }
    S->pending[w] = false;
  }
}

/// Copies `count` doubles of the device result buffer to the pinned mirror and
/// waits for the stream.
int fetch_results(slpb_solver* S, int count) {
  CU(cudaMemcpyAsync(S->h_results, S->d_results.p, count * sizeof(double),
                     cudaMemcpyDeviceToHost, S->stream));
  CU(cudaStreamSynchronize(S->stream));
  harvest_timers(S);
  S->counters.d2h_bytes += count * sizeof(double);
  return SLPB_OK;
}

CscView ae_view(slpb_solver* S) {
  return {S->ae_colptr.p, S->ae_rowidx.p, S->dvals.p + S->ad.off_ae};
}
CscView ai_view(slpb_solver* S) {
  return {S->ai_colptr.p, S->ai_rowidx.p, S->dvals.p + S->ad.off_ai};
}

/// Values (f, c_e, c_i) of the point whose leaves are in `leaf`.
int eval_values(slpb_solver* S, const double* leaf, double* vals) {
  CU(timer_begin(S, 1));
  int rc = run_sweep(S, S->pv, leaf, S->vstage.p);
  if (rc) return rc;
  rc = run_gather(S, S->gv, S->vstage.p, vals);
  if (rc) return rc;
  CU(timer_end(S, 1));
  ++S->counters.evals_values;
  return SLPB_OK;
}

int eval_derivs(slpb_solver* S, const double* leaf) {
  CU(timer_begin(S, 0));
  int rc = S->world > 1 ? run_sweep_sharded(S, S->pd, leaf, S->dstage.p)
                        : run_sweep(S, S->pd, leaf, S->dstage.p);
  if (rc) return rc;
  rc = run_gather(S, S->gd, S->dstage.p, S->dvals.p);
  if (rc) return rc;
  CU(timer_end(S, 0));
  ++S->counters.evals_full;
  return SLPB_OK;
}

int refresh_leaves(slpb_solver* S, const double* x, const double* y,
                   const double* z, double* leaf) {
  const int tot = S->n + S->me + S->mi;
  if (tot == 0) return SLPB_OK;
  k_prepare_leaves<<<blocks_for(tot, 256), 256, 0, S->stream>>>(
      x, y, z, S->d_c.p, S->n, S->me, S->mi, leaf);
  ++S->counters.kernel_launches;
  CU(cudaGetLastError());
  return SLPB_OK;
}

int point_info(slpb_solver* S, const double* vals, const double* s,
               double* d_out) {
  k_point_info<<<red_blocks(S->me + S->mi), kReduceThreads, 0, S->stream>>>(
      vals, s, S->me, S->mi, red_buf(S), d_out);
  ++S->counters.kernel_launches;
  CU(cudaGetLastError());
  return SLPB_OK;
}

void fill_point_info(const double* r, slpb_point_info* info) {
  info->f = r[0];
  info->ce_l1 = r[1];
  info->cis_l1 = r[2];
  info->log_s_sum = r[3];
  info->finite = static_cast<int32_t>(r[4]);
  info->ci_all_positive = static_cast<int32_t>(r[5]);
}

/// Where the merged calls park their results in the 64-double result block.
constexpr int kResStep = 0;    // 9 doubles of k_step_recover_stats
constexpr int kResPoint = 16;  // 6 (+1) doubles of k_point_info / k_deriv_finite
constexpr int kResKkt = 32;    // 25 doubles of k_kkt_stats

int kkt_stats_enqueue(slpb_solver* S, const double* c_e, const double* c_i,
                      const double* x, const double* s, const double* y,
                      const double* z, double mu, int offset,
                      bool with_deriv_finite = false) {
  k_kkt_stats<<<red_blocks(S->n + S->mi), kReduceThreads, 0, S->stream>>>(
      ae_view(S), ai_view(S), S->dvals.p + S->ad.off_g, c_e, c_i, x, s, y, z,
      S->d_c.p, S->inv_d_c.p, S->d_f, mu, S->n, S->me, S->mi, red_buf(S),
      S->d_results.p + offset, with_deriv_finite ? S->dvals.p : nullptr,
      S->ad.off_ae, S->ad.off_ai, S->ad.off_h, S->ad.off_h + S->ad.H.nnz(),
      with_deriv_finite ? S->d_results.p + kResPoint : nullptr);
  ++S->counters.kernel_launches;
  CU(cudaGetLastError());
  return SLPB_OK;
}

void fill_kkt_stats(const double* r, slpb_kkt_stats* out) {
  out->r_inf = r[0];
  out->r_l1 = r[1];
  out->y_l1 = r[2];
  out->z_l1 = r[3];
  out->sz_min = r[4];
  out->sz_max = r[5];
  out->sz_mu_l1 = r[6];
  out->ce_inf = r[7];
  out->ce_l1 = r[8];
  out->cis_inf = r[9];
  out->cis_l1 = r[10];
  out->u_r_inf = r[11];
  out->u_y_l1 = r[12];
  out->u_z_l1 = r[13];
  out->u_sz_min = r[14];
  out->u_sz_max = r[15];
  out->u_ce_inf = r[16];
  out->u_cis_inf = r[17];
  out->aetce_l2 = r[18];
  out->ce_l2 = r[19];
  out->aitcip_l2 = r[20];
  out->cip_l2 = r[21];
  out->x_inf = r[22];
  out->s_inf = r[23];
  out->xs_finite = static_cast<int32_t>(r[24]);
  out->pad = 0;
}

int kkt_stats(slpb_solver* S, const double* c_e, const double* c_i,
              const double* x, const double* s, const double* y,
              const double* z, double mu, slpb_kkt_stats* out) {
  int rc = kkt_stats_enqueue(S, c_e, c_i, x, s, y, z, mu, 0);
  if (rc) return rc;
  if ((rc = fetch_results(S, 25))) return rc;
  fill_kkt_stats(S->h_results, out);
  return SLPB_OK;
}

TreeView tree_view(slpb_solver* S) {
  TreeView T{};
  T.order = S->sy_level_supers.p;
  T.n_order = S->sym.n_super;
  T.order_bwd = S->sy_level_supers.p;
  T.n_fwd = S->sym.n_super;
  T.n_bwd = S->sym.n_super;
  T.metas = S->sy_metas.p;
  T.child_idx = S->sy_child_idx.p;
  T.rel_idx = S->sy_rel_idx.p;
  T.rows_idx = S->sy_rows_idx.p;
  T.asm_src = S->sy_asm_src.p;
  T.asm_dst = S->sy_asm_dst_ld.p;
  T.ext = ExtendList{S->sy_ext_src.p, S->sy_ext_dst.p};
  T.col_is_primal = S->sy_col_is_primal.p;
  T.perm = S->sy_perm.p;
  T.sync = S->tree_sync.p;
  T.n_super = S->sym.n_super;
  T.debug = S->tree_debug.p;
  return T;
}

/// Σ, S⁻¹, t and rhs = −[g − A_eᵀy − A_iᵀt ; c_e] (interior_point.hpp:444-448;
/// with the SOC accumulators, :611-616).
int build_rhs(slpb_solver* S, double mu, const double* cis_soc,
              const double* ce_for_rhs) {
  const int n = S->n, me = S->me, mi = S->mi;
  // Σ, S⁻¹, t and the right-hand side in one launch
  k_rhs_sigma<<<blocks_for(std::max(S->dim, mi), 256), 256, 0, S->stream>>>(
      ae_view(S), ai_view(S), S->dvals.p + S->ad.off_g, S->y.p, ce_for_rhs,
      S->s.p, S->z.p, S->vals_cur.p + 1 + me, cis_soc, mu, cis_soc ? 1 : 0, n,
      me, mi, S->sinv.p, S->sigma.p, S->tvec.p, S->rhs.p);
  ++S->counters.kernel_launches;
  CU(cudaGetLastError());
  return SLPB_OK;
}

/// Folds the finished event pairs of the collectives into comm_stats (after a
/// stream synchronisation).
void harvest_comm_timers(slpb_solver* S) {
  for (int w = 0; w < 3; ++w) {
    if (!S->cpending[w]) continue;
    float ms = 0.0f;
    if (cudaEventElapsedTime(&ms, S->cev[2 * w], S->cev[2 * w + 1]) ==
        cudaSuccess) {
      if (!S->comm_warm[w]) {
        // the first collective of a kind (a new message size) pays NCCL's
        // one-time buffer / protocol set-up: not part of the steady state
        S->comm_warm[w] = true;
        --S->comm_timed[w];
      } else {
        S->comm_stats.total_ms[w] += ms;
      }
    } else {
      cudaGetLastError();
    }
    S->cpending[w] = false;
  }
}

int nccl_all_gather(slpb_solver* S, const double* mine, double* all,
                    size_t count) {
  const int rc =
      nccl_api().AllGather(mine, all, count, kNcclFloat64, S->comm, S->stream);
  if (rc != 0) {
    return fail(S, SLPB_ERR_NCCL,
                std::string("ncclAllGather: ") +
                    (nccl_api().GetErrorString ? nccl_api().GetErrorString(rc)
                                               : "error"));
  }
  return SLPB_OK;
}

/// Sharded factorisation: every rank sends the update matrices and update
/// vectors of its subtree roots (and, with_stats, its inertia counts) and
/// receives everybody else's — ONE all-gather of a few KB per rank.
int exchange_roots(slpb_solver* S, int nv, bool with_stats) {
  const Symbolic& Y = S->sym;
  const TreeShard& H = S->tshard;
  const int W = H.world, r = S->rank;
  const int seg = kPackStats + S->ts_pack_max;
  // (timed only when the previous event pair has been read back: no host
  // synchronisation is added for the sake of the statistics)
  const bool timed = !S->cpending[1];
  if (timed) CU(cudaEventRecord(S->cev[2], S->stream));
  double* mine = S->ts_pack_buf.p + size_t(r) * nv * seg;
  const int len = static_cast<int>(H.rank_roots.empty() ? 0 : S->ts_pack_len_host[r]);
  const dim3 gp(blocks_for(std::max(len, kPackStats), 256), nv);
  k_tree_pack<<<gp, 256, 0, S->stream>>>(
      S->ts_pack_pos.p + size_t(r) * S->ts_pack_max, len, seg, nv,
      S->updates.p, Y.update_size, S->uvecs.p,
      static_cast<int64_t>(Y.rel_ptr.back()), S->fstats.p, mine);
  int rc = nccl_all_gather(S, mine, S->ts_pack_buf.p, size_t(nv) * seg);
  if (rc) return rc;
  const dim3 gu(blocks_for(std::max(S->ts_pack_max, 1), 256), W * nv);
  k_tree_unpack<<<gu, 256, 0, S->stream>>>(
      S->ts_pack_pos.p, S->ts_pack_len.p, S->ts_pack_max, seg, nv, W, r,
      S->ts_pack_buf.p, S->updates.p, Y.update_size, S->uvecs.p,
      static_cast<int64_t>(Y.rel_ptr.back()), S->fstats.p, with_stats ? 1 : 0);
  if (timed) {
    CU(cudaEventRecord(S->cev[3], S->stream));
    S->cpending[1] = true;
    ++S->comm_timed[1];
  }
  S->counters.kernel_launches += 2;
  ++S->comm_stats.count[1];
  S->comm_stats.bytes[1] += int64_t(nv) * seg * 8;
  CU(cudaGetLastError());
  return SLPB_OK;
}

/// Sharded solve: the solution entries of every rank's own columns travel to
/// all the others (the top's are computed everywhere).
int exchange_solution(slpb_solver* S) {
  const TreeShard& H = S->tshard;
  const int W = H.world, r = S->rank;
  const bool timed = !S->cpending[2];
  if (timed) CU(cudaEventRecord(S->cev[4], S->stream));
  double* mine = S->ts_sol_buf.p + size_t(r) * S->ts_sol_max;
  const int len = S->ts_sol_len_host[r];
  if (len > 0) {
    k_gather_idx<<<blocks_for(len, 256), 256, 0, S->stream>>>(
        S->sol.p, S->ts_sol_idx.p + size_t(r) * S->ts_sol_max, len, mine);
  }
  int rc = nccl_all_gather(S, mine, S->ts_sol_buf.p, S->ts_sol_max);
  if (rc) return rc;
  const dim3 g(blocks_for(S->ts_sol_max, 256), W);
  k_scatter_idx<<<g, 256, 0, S->stream>>>(S->ts_sol_buf.p, S->ts_sol_idx.p,
                                          S->ts_sol_len.p, S->ts_sol_max, W, r,
                                          S->sol.p);
  if (timed) {
    CU(cudaEventRecord(S->cev[5], S->stream));
    S->cpending[2] = true;
    ++S->comm_timed[2];
  }
  S->counters.kernel_launches += 2;
  ++S->comm_stats.count[2];
  S->comm_stats.bytes[2] += int64_t(S->ts_sol_max) * 8;
  CU(cudaGetLastError());
  return SLPB_OK;
}

int launch_solve(slpb_solver* S, bool skip_forward) {
  const Symbolic& Y = S->sym;
  if (S->group != nullptr) {
    CU(cudaStreamSynchronize(S->stream));  // the rhs is complete
    harvest_timers(S);
    return group_solve(S);  // forward + backward in the batched launch → S->sol
  }
  CU(timer_begin(S, 4));
  if (S->use_tree) {
    const int sel = S->factor_sel;
    const double* panels = S->panels.p + sel * Y.panel_size;
    const double* Dsel = S->D.p + size_t(sel) * Y.dim;
    double* xperm = S->xperm.p + size_t(sel) * Y.dim;
    double* uvecs = S->uvecs.p + size_t(sel) * Y.rel_ptr.back();
    TreeView T = tree_view(S);
    T.sync = S->solve_sync.p;  // the solve's own dependency words
    if (S->hybrid) {
      // forward: the small fronts in one tree launch, the top level by level;
      // backward: the top level by level, then the small fronts (whose parents
      // in the top count as done: hy_bflag_init)
      const int top_smem = Y.max_front * static_cast<int>(sizeof(double));
      const int n_top_levels = static_cast<int>(S->hy_level_off.size()) - 1;
      T.order = S->hy_small_order.p;
      T.order_bwd = S->hy_small_order.p;
      k_init_solve_sync<<<blocks_for(Y.n_super, 256), 256, 0, S->stream>>>(
          S->solve_sync.p, Y.n_super, 0, nullptr, nullptr);
      T.n_fwd = S->n_small;
      T.n_bwd = 0;
      k_solve_tree<<<S->solve_blocks, kTreeWarps * 32, 0, S->stream>>>(
          T, panels, Dsel, S->rhs.p, xperm, uvecs, S->sol.p);
      for (int L = 0; L < n_top_levels; ++L) {
        const int cnt = S->hy_level_off[L + 1] - S->hy_level_off[L];
        k_forward_level<<<cnt, kFrontThreads, top_smem, S->stream>>>(
            S->sview, S->hy_top_supers.p + S->hy_level_off[L], panels, S->rhs.p,
            xperm, uvecs);
      }
      for (int L = n_top_levels - 1; L >= 0; --L) {
        const int cnt = S->hy_level_off[L + 1] - S->hy_level_off[L];
        k_backward_level<<<cnt, kFrontThreads, top_smem, S->stream>>>(
            S->sview, S->hy_top_supers.p + S->hy_level_off[L], panels, Dsel,
            xperm);
      }
      k_init_solve_sync<<<blocks_for(Y.n_super, 256), 256, 0, S->stream>>>(
          S->solve_sync.p, Y.n_super, 1, nullptr, S->hy_bflag_init.p);
      T.n_fwd = 0;
      T.n_bwd = S->n_small;
      k_solve_tree<<<S->solve_blocks, kTreeWarps * 32, 0, S->stream>>>(
          T, panels, Dsel, S->rhs.p, xperm, uvecs, S->sol.p);
      // (the tree kernel un-permutes its own fronts; this covers the top)
      k_unpermute<<<blocks_for(S->dim, 256), 256, 0, S->stream>>>(
          xperm, S->sy_perm.p, S->dim, S->sol.p);
      S->counters.kernel_launches += 5 + 2 * n_top_levels;
      S->solve_sync_preset = false;
    } else if (!S->tree_sharded) {
      if (!(skip_forward && S->solve_sync_preset)) {
        k_init_solve_sync<<<blocks_for(Y.n_super, 256), 256, 0, S->stream>>>(
            S->solve_sync.p, Y.n_super, skip_forward ? 1 : 0, nullptr);
        ++S->counters.kernel_launches;
      }
      S->solve_sync_preset = false;  // the words are used up
      if (skip_forward) T.n_fwd = 0;
      k_solve_tree<<<S->solve_blocks, kTreeWarps * 32, 0, S->stream>>>(
          T, panels, Dsel, S->rhs.p, xperm, uvecs, S->sol.p);
      ++S->counters.kernel_launches;
    } else {
      // Sharded: forward over the own subtrees, exchange of the roots' update
      // vectors, then forward over the replicated top and backward over top
      // and own subtrees; the solution pieces travel last.
      const TreeShard& H = S->tshard;
      const int n_mine = static_cast<int>(H.rank_order[S->rank].size());
      const int n_top = static_cast<int>(H.top_order.size());
      if (!skip_forward) {
        k_init_solve_sync<<<blocks_for(Y.n_super, 256), 256, 0, S->stream>>>(
            S->solve_sync.p, Y.n_super, 0, nullptr);
        T.order = S->ts_my_order.p;
        T.n_fwd = n_mine;
        T.n_bwd = 0;
        if (n_mine > 0) {
          k_solve_tree<<<S->solve_blocks, kTreeWarps * 32, 0, S->stream>>>(
              T, panels, Dsel, S->rhs.p, xperm, uvecs, S->sol.p);
        }
        S->counters.kernel_launches += 2;
        // (the roots' data of the selected variant sits at variant offset sel:
        // exchange both halves' worth only when both exist)
        int rc = exchange_roots(S, sel + 1, false);
        if (rc) return rc;
      }
      k_init_solve_sync<<<blocks_for(Y.n_super, 256), 256, 0, S->stream>>>(
          S->solve_sync.p, Y.n_super, skip_forward ? 1 : 0,
          S->ts_top_fcount_init.p);
      T.order = S->ts_top_order.p;
      T.n_fwd = skip_forward ? 0 : n_top;
      T.order_bwd = S->ts_bwd_order.p;
      T.n_bwd = n_mine + n_top;
      k_solve_tree<<<S->solve_blocks, kTreeWarps * 32, 0, S->stream>>>(
          T, panels, Dsel, S->rhs.p, xperm, uvecs, S->sol.p);
      S->counters.kernel_launches += 2;
      int rc = exchange_solution(S);
      if (rc) return rc;
    }
  } else {
    const int smem = Y.max_front * static_cast<int>(sizeof(double));
    const double* panels = S->panels.p + S->factor_sel * Y.panel_size;
    const double* Dsel = S->D.p + size_t(S->factor_sel) * Y.dim;
    for (int L = 0; L < Y.n_levels; ++L) {
      const int cnt = Y.level_ptr[L + 1] - Y.level_ptr[L];
      k_forward_level<<<cnt, kFrontThreads, smem, S->stream>>>(
          S->sview, S->sy_level_supers.p + Y.level_ptr[L], panels, S->rhs.p,
          S->xperm.p, S->uvecs.p);
    }
    for (int L = Y.n_levels - 1; L >= 0; --L) {
      const int cnt = Y.level_ptr[L + 1] - Y.level_ptr[L];
      k_backward_level<<<cnt, kFrontThreads, smem, S->stream>>>(
          S->sview, S->sy_level_supers.p + Y.level_ptr[L], panels, Dsel,
          S->xperm.p);
    }
    k_unpermute<<<blocks_for(S->dim, 256), 256, 0, S->stream>>>(
        S->xperm.p, S->sy_perm.p, S->dim, S->sol.p);
    S->counters.kernel_launches += 2 * Y.n_levels + 1;
  }
  CU(timer_end(S, 4));
  CU(cudaGetLastError());
  ++S->counters.solves;
  return SLPB_OK;
}

void fill_step_info(const double* r, slpb_step_info* info);

/// rhs → solve → step recovery → step stats, into the given step arrays.
int solve_into(slpb_solver* S, double mu, double tau, bool soc,
               const double* cis_soc, const double* ce_for_rhs, double* px,
               double* ps, double* py, double* pz, slpb_step_info* info) {
  const int n = S->n, me = S->me, mi = S->mi;
  // the factorisation already carried this right-hand side forward?
  // (`soc` and not `cis_soc != nullptr`: without inequality constraints the
  // corrected c_i − s array is empty and its pointer null)
  const bool fused = !soc && S->use_tree && S->rhs_ready &&
                     S->rhs_mu == mu && S->fwd_valid[S->factor_sel];
  if (!fused) {
    int rc0 = build_rhs(S, mu, cis_soc, ce_for_rhs);
    if (rc0) return rc0;
    // a different right-hand side now sits in rhs/xperm/uvecs
    S->rhs_ready = false;
    S->fwd_valid[0] = S->fwd_valid[1] = false;
  }
  int rc = launch_solve(S, fused);
  if (rc) return rc;
  // step recovery and its reductions in one launch
  k_step_recover_stats<<<red_blocks(n + mi), kReduceThreads, 0, S->stream>>>(
      S->sol.p, S->ai_rowptr.p, S->ai_rcol.p, S->ai_ridx.p,
      S->dvals.p + S->ad.off_ai, S->vals_cur.p + 1 + me, S->s.p, S->z.p,
      cis_soc, S->sinv.p, S->sigma.p, S->dvals.p + S->ad.off_g, mu, tau, n, me,
      mi, px, py, ps, pz, red_buf(S), S->d_results.p);
  ++S->counters.kernel_launches;
  CU(cudaGetLastError());
  if (info == nullptr) return SLPB_OK;  // the caller fetches (slpb_solve_trial)
  rc = fetch_results(S, 9);
  if (rc) return rc;
  fill_step_info(S->h_results, info);
  return SLPB_OK;
}

void fill_step_info(const double* r, slpb_step_info* info) {
  info->alpha_max = r[0];
  info->alpha_z = r[1];
  info->g_dot_px = r[2];
  info->sinv_dot_ps = r[3];
  info->px_inf = r[4];
  info->ps_inf = r[5];
  info->py_inf = r[6];
  info->pz_inf = r[7];
  info->finite = static_cast<int32_t>(r[8]);
  info->pad = 0;
}

}  // namespace slpb

// ===========================================================================
// C ABI
// ===========================================================================

extern "C" {

int slpb_create(int device, slpb_solver** out) {
  if (!out) return SLPB_ERR_ARGUMENT;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
    return SLPB_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= count) return SLPB_ERR_ARGUMENT;
  if (cudaSetDevice(device) != cudaSuccess) return SLPB_ERR_CUDA;
  auto S = std::make_unique<slpb_solver>();
  S->device = device;
  if (const char* mode = std::getenv("SLPB_FACTOR_ARITH")) {
    if (std::strcmp(mode, "tensor") == 0) S->factor_arith = SLPB_ARITH_TENSOR;
  }
  if (cudaStreamCreateWithFlags(&S->stream, cudaStreamNonBlocking) !=
      cudaSuccess) {
    return SLPB_ERR_CUDA;
  }
  {
    // keep freed memory in the device's default pool (re-used by the next
    // solve) instead of returning it to the driver at every synchronisation
    cudaMemPool_t pool = nullptr;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    } else {
      cudaGetLastError();
    }
  }
  const AllocScope alloc_scope{S->stream};
  S->h_results = pinned_pool().acquire();
  if (!S->h_results) return SLPB_ERR_CUDA;
  if (S->d_results.alloc(kResultDoubles) != cudaSuccess) return SLPB_ERR_CUDA;
  if (S->red_partials.alloc(kReduceBlocks * 32) != cudaSuccess ||
      S->red_counter.alloc(1) != cudaSuccess ||
      S->red_counter.zero(S->stream) != cudaSuccess) {
    return SLPB_ERR_CUDA;
  }
  for (auto& e : S->ev) {
    if (cudaEventCreate(&e) != cudaSuccess) return SLPB_ERR_CUDA;
  }
  for (int i = 0; i < slpb_solver::kSideStreams; ++i) {
    if (cudaStreamCreateWithFlags(&S->side[i], cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&S->join_ev[i], cudaEventDisableTiming) != cudaSuccess) {
      return SLPB_ERR_CUDA;
    }
  }
  if (cudaEventCreateWithFlags(&S->fork_ev, cudaEventDisableTiming) != cudaSuccess) {
    return SLPB_ERR_CUDA;
  }
  *out = S.release();
  return SLPB_OK;
}

void slpb_destroy(slpb_solver* S) {
  if (!S) return;
  if (S->group) slpb_group_leave(S);
  const bool timing = std::getenv("SLPB_DESTROY_TIMING") != nullptr;
  const auto t_begin = std::chrono::steady_clock::now();
  auto lap = [&, last = t_begin](const char* what) mutable {
    if (!timing) return;
    const auto now = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[slpb destroy] %-22s %8.2f ms\n", what,
                 std::chrono::duration<double, std::milli>(now - last).count());
    last = now;
  };
  cudaSetDevice(S->device);
  if (S->stream) cudaStreamSynchronize(S->stream);
  lap("stream sync");
  for (auto& e : S->ev) {
    if (e) cudaEventDestroy(e);
  }
  for (auto& e : S->cev) {
    if (e) cudaEventDestroy(e);
  }
  for (int i = 0; i < slpb_solver::kSideStreams; ++i) {
    if (S->side[i]) {
      cudaStreamSynchronize(S->side[i]);
      cudaStreamDestroy(S->side[i]);
    }
    if (S->join_ev[i]) cudaEventDestroy(S->join_ev[i]);
  }
  if (S->fork_ev) cudaEventDestroy(S->fork_ev);
  lap("events");
  if (S->comm && nccl_api().ok) nccl_api().CommDestroy(S->comm);
  if (S->h_results) pinned_pool().give_back(S->h_results);
  lap("comm + pinned");
  cudaStream_t st = S->stream;
  {
    const AllocScope alloc_scope{st};
    delete S;  // DevBufs are freed stream-ordered; the stream outlives them
  }
  lap("delete (async frees)");
  if (st) cudaStreamDestroy(st);
  lap("stream destroy");
}

int slpb_comm_unique_id(void* id_out) {
  if (!id_out) return SLPB_ERR_ARGUMENT;
  if (!nccl_api().ok) return SLPB_ERR_NCCL;
  NcclUniqueId id;
  if (nccl_api().GetUniqueId(&id) != 0) return SLPB_ERR_NCCL;
  std::memcpy(id_out, &id, sizeof(id));
  return SLPB_OK;
}

int slpb_comm_init(slpb_solver* S, int rank, int world, const void* id) {
  if (!S || !id || world < 1 || rank < 0 || rank >= world) {
    return SLPB_ERR_ARGUMENT;
  }
  if (S->finalized) {
    return fail(S, SLPB_ERR_STATE, "slpb_comm_init must precede slpb_finalize");
  }
  if (world == 1) return SLPB_OK;
  if (!nccl_api().ok) {
    return fail(S, SLPB_ERR_NCCL, "libnccl.so.2 could not be loaded");
  }
  CU(cudaSetDevice(S->device));
  NcclUniqueId uid;
  std::memcpy(&uid, id, sizeof(uid));
  const int rc = nccl_api().CommInitRank(&S->comm, world, uid, rank);
  if (rc != 0) {
    return fail(S, SLPB_ERR_NCCL,
                std::string("ncclCommInitRank: ") +
                    (nccl_api().GetErrorString ? nccl_api().GetErrorString(rc)
                                               : "error"));
  }
  S->rank = rank;
  S->world = world;
  for (auto& e : S->cev) {
    if (!e) CU(cudaEventCreate(&e));
  }
  // NCCL sets its channels up lazily inside the first collective (seconds):
  // pay for that here, not inside the first Newton iteration
  int32_t any = 0;
  return slpb_comm_agree(S, 0, &any);
}

int slpb_comm_agree(slpb_solver* S, int32_t local_flag, int32_t* any) {
  if (!S || !any) return SLPB_ERR_ARGUMENT;
  *any = local_flag != 0;
  if (S->world <= 1) return SLPB_OK;
  CU(cudaSetDevice(S->device));
  const AllocScope alloc_scope{S->stream};
  if (S->agree_buf.n < size_t(S->world)) CU(S->agree_buf.alloc(S->world));
  const double mine = local_flag != 0 ? 1.0 : 0.0;
  CU(cudaMemcpyAsync(S->agree_buf.p + S->rank, &mine, sizeof(double),
                     cudaMemcpyHostToDevice, S->stream));
  const int rc = nccl_api().AllGather(S->agree_buf.p + S->rank, S->agree_buf.p,
                                      1, kNcclFloat64, S->comm, S->stream);
  if (rc != 0) return fail(S, SLPB_ERR_NCCL, "ncclAllGather (slpb_comm_agree)");
  std::vector<double> all(S->world, 0.0);
  CU(cudaMemcpyAsync(all.data(), S->agree_buf.p, S->world * sizeof(double),
                     cudaMemcpyDeviceToHost, S->stream));
  CU(cudaStreamSynchronize(S->stream));
  for (double v : all) *any |= v != 0.0;
  return SLPB_OK;
}

const char* slpb_last_error(const slpb_solver* S) {
  return S ? S->error.c_str() : "null handle";
}

int slpb_upload_tape(slpb_solver* S, int32_t n_nodes, const uint8_t* op,
                     const int32_t* lhs, const int32_t* rhs, const double* val,
                     int32_t n_x, const int32_t* leaf_x, int32_t n_y,
                     const int32_t* leaf_y, int32_t n_z,
                     const int32_t* leaf_z) {
  if (!S) return SLPB_ERR_ARGUMENT;
  if (!ingest_tape(S->tape, n_nodes, op, lhs, rhs, val, n_x, leaf_x, n_y,
                   leaf_y, n_z, leaf_z, S->error)) {
    return SLPB_ERR_ARGUMENT;
  }
  S->n = n_x;
  S->me = n_y;
  S->mi = n_z;
  S->dim = n_x + n_y;
  S->have_tape = true;
  S->finalized = S->analyzed = false;
  for (auto& r : S->rows) r = RowSet{};
  S->counters.tape_nodes = n_nodes;
  S->counters.h2d_bytes += int64_t(n_nodes) * 17;
  return SLPB_OK;
}

int slpb_upload_rows(slpb_solver* S, int which, const slpb_rowset* rows,
                     const double* const_val) {
  if (!S || which < 0 || which >= SLPB_OUT_COUNT) return SLPB_ERR_ARGUMENT;
  if (!S->have_tape) {
    return fail(S, SLPB_ERR_STATE, "slpb_upload_rows before slpb_upload_tape");
  }
  if (!ingest_rows(S->rows[which], S->tape, which, rows, const_val, S->error)) {
    return SLPB_ERR_ARGUMENT;
  }
  return SLPB_OK;
}

int slpb_set_ignore_constraint_hessian(slpb_solver* S, int ignore) {
  if (!S) return SLPB_ERR_ARGUMENT;
  if (S->finalized) {
    return fail(S, SLPB_ERR_STATE,
                "slpb_set_ignore_constraint_hessian must precede slpb_finalize");
  }
  S->ignore_h_c = ignore != 0;
  return SLPB_OK;
}

int slpb_set_factor_arithmetic(slpb_solver* S, int mode) {
  if (!S) return SLPB_ERR_ARGUMENT;
  if (mode != SLPB_ARITH_REFERENCE && mode != SLPB_ARITH_TENSOR) {
    return fail(S, SLPB_ERR_ARGUMENT, "unknown factor arithmetic mode");
  }
  S->factor_arith = mode;
  return SLPB_OK;
}

int slpb_finalize(slpb_solver* S) {
  if (!S) return SLPB_ERR_ARGUMENT;
  const AllocScope alloc_scope{S->stream};
  if (!S->have_tape) {
    return fail(S, SLPB_ERR_STATE, "slpb_finalize before slpb_upload_tape");
  }
  CU(cudaSetDevice(S->device));
  // The symbolic analysis of the KKT system needs only the static patterns,
  // which the compiler has after its first step: the assembly recipe and the
  // analysis in the default ordering run on another thread while the programs
  // are being built (slpb_analyze picks the result up).
  S->sym_ahead_ok = false;
  std::future<void> ahead;
  const auto start_analysis = [&]() {
    ahead = std::async(std::launch::async, [S]() {
      build_kkt_recipe(S->n, S->me, S->ad.H, S->ad.A_e, S->ad.A_i, S->recipe);
      try {
        std::string err;
        S->sym_ahead_ok =
            analyze_kkt(S->recipe.K, S->n, SLPB_ORDER_NESTED_DISSECTION, nullptr,
                        S->sym_ahead, err);
      } catch (...) {
        S->sym_ahead_ok = false;  // slpb_analyze will run it again and report
      }
    });
  };
  const bool compiled =
      compile_autodiff(S->tape, S->rows, S->ignore_h_c, S->ad, start_analysis);
  try {
    if (ahead.valid()) ahead.get();
  } catch (const std::exception& e) {
    return fail(S, SLPB_ERR_UNSUPPORTED,
                std::string("building the KKT assembly recipe: ") + e.what());
  }
  if (!compiled) {
    S->error = S->ad.error;
    return SLPB_ERR_UNSUPPORTED;
  }
  const int n = S->n, me = S->me, mi = S->mi;
  int rc;
  if ((rc = upload_program_set(S, S->ad.values, S->pv))) return rc;
  if ((rc = upload_program_set(S, S->ad.derivs, S->pd))) return rc;
  if ((rc = upload_gather(S, S->ad.value_gather, S->gv))) return rc;
  if ((rc = upload_gather(S, S->ad.deriv_gather, S->gd))) return rc;
  CU(S->vstage.upload(S->ad.value_stage_init, S->stream));
  CU(S->dstage.upload(S->ad.deriv_stage_init, S->stream));
  if (S->world > 1) {
    build_shard_plan(S->ad.derivs, S->world, S->shard);
    CU(S->shard_slots.upload(S->shard.slots, S->stream));
    CU(S->shard_buf.alloc(std::max<size_t>(1, S->shard.slots.size())));
  }
  // Function attributes are per process, not per handle: only ever raise them
  // (a second handle — feasibility restoration, multistart — must not lower
  // what a live handle relies on).
  CU(raise_dynamic_smem(k_ad_sweep, std::max(S->pv.max_smem, S->pd.max_smem)));

  CU(S->leaf_cur.alloc(n + me + mi));
  CU(S->leaf_trial.alloc(n + me + mi));
  CU(S->leaf_cur.zero(S->stream));
  CU(S->leaf_trial.zero(S->stream));
  std::vector<double> ones(me + mi, 1.0);
  CU(S->d_c.upload(ones, S->stream));
  CU(S->inv_d_c.upload(ones, S->stream));
  S->d_f = 1.0;
  for (auto* b : {&S->x, &S->tx, &S->px, &S->spx}) CU(b->alloc(n));
  for (auto* b : {&S->y, &S->ty, &S->py, &S->spy, &S->ce_soc}) CU(b->alloc(me));
  for (auto* b : {&S->s, &S->z, &S->ts, &S->tz, &S->ps, &S->pz, &S->sps,
                  &S->spz, &S->cis_soc, &S->sigma, &S->sinv, &S->tvec}) {
    CU(b->alloc(mi));
  }
  CU(S->vals_cur.alloc(1 + me + mi));
  CU(S->vals_trial.alloc(1 + me + mi));
  CU(S->vals_cur.zero(S->stream));
  CU(S->vals_trial.zero(S->stream));
  CU(S->dvals.alloc(S->ad.off_h + S->ad.H.nnz()));
  CU(S->dvals.zero(S->stream));
  CU(S->ae_colptr.upload(S->ad.A_e.colptr, S->stream));
  CU(S->ae_rowidx.upload(S->ad.A_e.rowidx, S->stream));
  CU(S->ai_colptr.upload(S->ad.A_i.colptr, S->stream));
  CU(S->ai_rowidx.upload(S->ad.A_i.rowidx, S->stream));
  {
    const Pattern& A = S->ad.A_i;
    std::vector<int32_t> rptr(mi + 1, 0), rcol(A.nnz()), ridx(A.nnz());
    for (int32_t r : A.rowidx) ++rptr[r + 1];
    for (int i = 0; i < mi; ++i) rptr[i + 1] += rptr[i];
    std::vector<int32_t> nxt(rptr.begin(), rptr.end() - 1);
    for (int32_t c = 0; c < A.cols; ++c) {
      for (int32_t k = A.colptr[c]; k < A.colptr[c + 1]; ++k) {
        const int32_t q = nxt[A.rowidx[k]]++;
        rcol[q] = c;
        ridx[q] = k;
      }
    }
    CU(S->ai_rowptr.upload(rptr, S->stream));
    CU(S->ai_rcol.upload(rcol, S->stream));
    CU(S->ai_ridx.upload(ridx, S->stream));
  }
  // KKT recipe (built beside the compiler, see above)
  CU(S->k_h_idx.upload(S->recipe.h_idx, S->stream));
  CU(S->k_ae_idx.upload(S->recipe.ae_idx, S->stream));
  CU(S->k_prod_ptr.upload(S->recipe.prod_ptr, S->stream));
  CU(S->k_prod_a.upload(S->recipe.prod_a, S->stream));
  CU(S->k_prod_b.upload(S->recipe.prod_b, S->stream));
  CU(S->k_prod_row.upload(S->recipe.prod_row, S->stream));
  {
    std::vector<uint8_t> flag(S->recipe.K.nnz(), 0);
    for (int c = 0; c < n; ++c) flag[S->recipe.diag_idx[c]] = 1;
    CU(S->k_diag_flag.upload(flag, S->stream));
  }
  CU(S->Kval.alloc(S->recipe.K.nnz()));
  CU(S->rhs.alloc(S->dim));
  CU(S->sol.alloc(S->dim));
  CU(cudaStreamSynchronize(S->stream));
  S->counters.program_bytes =
      int64_t(S->ad.values.blob.size() + S->ad.derivs.blob.size() +
              S->ad.values.task_bindings.size() +
              S->ad.derivs.task_bindings.size()) * 4;
  S->counters.n_clusters = int64_t(S->ad.values.cluster_prog.size() +
                                   S->ad.derivs.cluster_prog.size());
  S->counters.n_program_classes = int64_t(S->ad.values.prog_offset.size() +
                                          S->ad.derivs.prog_offset.size());
  S->counters.h2d_bytes += S->counters.program_bytes;
  S->finalized = true;
  S->analyzed = false;
  return SLPB_OK;
}

int slpb_set_scaling(slpb_solver* S, double d_f, const double* d_ce,
                     const double* d_ci) {
  if (!S || !S->finalized) return SLPB_ERR_STATE;
  if ((S->me > 0 && !d_ce) || (S->mi > 0 && !d_ci)) return SLPB_ERR_ARGUMENT;
  CU(cudaSetDevice(S->device));
  S->d_f = d_f;
  std::vector<double> d(S->me + S->mi);
  std::copy(d_ce, d_ce + S->me, d.begin());
  std::copy(d_ci, d_ci + S->mi, d.begin() + S->me);
  if (!d.empty()) {
    CU(cudaMemcpyAsync(S->d_c.p, d.data(), d.size() * sizeof(double),
                       cudaMemcpyHostToDevice, S->stream));
    std::vector<double> inv(d.size());
    for (size_t i = 0; i < d.size(); ++i) inv[i] = 1.0 / d[i];
    CU(cudaMemcpyAsync(S->inv_d_c.p, inv.data(), inv.size() * sizeof(double),
                       cudaMemcpyHostToDevice, S->stream));
    CU(cudaStreamSynchronize(S->stream));
  }
  return SLPB_OK;
}

int slpb_analyze(slpb_solver* S, int ordering, const int32_t* perm,
                 slpb_symbolic_stats* stats) {
  if (!S || !S->finalized) return SLPB_ERR_STATE;
  const AllocScope alloc_scope{S->stream};
  CU(cudaSetDevice(S->device));
  if (ordering == SLPB_ORDER_NESTED_DISSECTION && perm == nullptr &&
      S->sym_ahead_ok) {
    S->sym = std::move(S->sym_ahead);  // analysed beside the compiler
    S->sym_ahead_ok = false;
  } else if (!analyze_kkt(S->recipe.K, S->n, ordering, perm, S->sym, S->error)) {
    return SLPB_ERR_ARGUMENT;
  }
  const Symbolic& Y = S->sym;
  const size_t front_smem =
      (size_t(Y.max_front) * Y.max_front + Y.max_front) * sizeof(double);
  // One block per front keeps the front in shared memory up to 200 KB (order
  // 158); a bigger front — a variable shared by hundreds of stages — gets a
  // workspace in global memory instead: slots for as many of them as one level
  // of the tree holds.
  constexpr size_t kFrontSmemCap = 200 * 1024;
  S->front_cap_doubles = static_cast<int>(kFrontSmemCap / sizeof(double));
  S->front_smem_bytes = static_cast<int>(std::min(front_smem, kFrontSmemCap));
  S->front_stride = int64_t(Y.max_front) * Y.max_front + Y.max_front;
  {
    std::vector<int32_t> slot(Y.n_super, 0);
    int32_t max_slots = 0;
    for (int L = 0; L < Y.n_levels; ++L) {
      int32_t used = 0;
      for (int k = Y.level_ptr[L]; k < Y.level_ptr[L + 1]; ++k) {
        const int32_t q = Y.level_supers[k];
        const int64_t F = Y.front_dim[q];
        if (F * F + F > S->front_cap_doubles) slot[q] = used++;
      }
      max_slots = std::max(max_slots, used);
    }
    if (max_slots > 0) {
      if (size_t(max_slots) * S->front_stride * sizeof(double) > (size_t(8) << 30)) {
        return fail(S, SLPB_ERR_UNSUPPORTED,
                    "the frontal matrices of one level exceed 8 GB (front order " +
                        std::to_string(Y.max_front) + ")");
      }
      CU(S->front_scratch.alloc(size_t(max_slots) * S->front_stride));
    }
    CU(S->front_slot.upload(slot, S->stream));
  }
  CU(raise_dynamic_smem(k_factor_level, S->front_smem_bytes));
  CU(S->sy_super_first.upload(Y.super_first, S->stream));
  CU(S->sy_front_dim.upload(Y.front_dim, S->stream));
  CU(S->sy_rows_ptr.upload(Y.rows_ptr, S->stream));
  CU(S->sy_rows_idx.upload(Y.rows_idx, S->stream));
  CU(S->sy_panel_ptr.upload(Y.panel_ptr, S->stream));
  CU(S->sy_update_ptr.upload(Y.update_ptr, S->stream));
  CU(S->sy_child_ptr.upload(Y.child_ptr, S->stream));
  CU(S->sy_child_idx.upload(Y.child_idx, S->stream));
  CU(S->sy_rel_ptr.upload(Y.rel_ptr, S->stream));
  CU(S->sy_rel_idx.upload(Y.rel_idx, S->stream));
  CU(S->sy_asm_ptr.upload(Y.asm_ptr, S->stream));
  CU(S->sy_asm_src.upload(Y.asm_src, S->stream));
  CU(S->sy_asm_dst.upload(Y.asm_dst, S->stream));
  CU(S->sy_col_is_primal.upload(Y.col_is_primal, S->stream));
  CU(S->sy_perm.upload(Y.perm, S->stream));
  CU(S->sy_level_supers.upload(Y.level_supers, S->stream));
  // two variants: slpb_factor_pair factors two regularisations side by side
  CU(S->panels.alloc(2 * Y.panel_size));
  CU(S->updates.alloc(2 * Y.update_size));
  CU(S->D.alloc(2 * size_t(Y.dim)));
  CU(S->uvecs.alloc(2 * size_t(Y.rel_ptr.back())));
  CU(S->xperm.alloc(2 * size_t(Y.dim)));
  CU(S->fstats.alloc(16));
  // fronts of order ≤ 32: one warp per front, one launch per factorisation;
  // larger ones and their ancestors: the level kernels (hybrid)
  std::vector<uint8_t> in_top(Y.n_super, 0);
  S->hybrid = false;
  for (int32_t s : Y.level_supers) {  // ascending level: children come first
    if (Y.front_dim[s] > 32) in_top[s] = 1;
    if (in_top[s]) {
      S->hybrid = true;
      if (Y.super_parent[s] >= 0) in_top[Y.super_parent[s]] = 1;
    }
  }
  S->n_small = 0;
  S->hy_level_off.clear();
  if (S->hybrid) {
    std::vector<int32_t> small, top, bflag(Y.n_super, 0);
    S->hy_level_off.push_back(0);
    for (int L = 0; L < Y.n_levels; ++L) {
      for (int k = Y.level_ptr[L]; k < Y.level_ptr[L + 1]; ++k) {
        const int32_t s = Y.level_supers[k];
        (in_top[s] ? top : small).push_back(s);
        bflag[s] = in_top[s];
      }
      if (static_cast<int32_t>(top.size()) > S->hy_level_off.back()) {
        S->hy_level_off.push_back(static_cast<int32_t>(top.size()));
      }
    }
    S->n_small = static_cast<int>(small.size());
    if (small.empty()) small.push_back(0);
    CU(S->hy_small_order.upload(small, S->stream));
    CU(S->hy_top_supers.upload(top, S->stream));
    CU(S->hy_bflag_init.upload(bflag, S->stream));
  }
  S->use_tree = !S->hybrid || S->n_small > 0;
  if (S->hybrid && std::getenv("SLPB_NO_HYBRID")) {
    // development switch: round 1's behaviour (every front through the level
    // kernels as soon as one front exceeds order 32)
    S->hybrid = false;
    S->use_tree = false;
  }
  S->tree_smem_doubles = kFrontSmemDoubles;
  if (S->use_tree) {
    // the warp kernels keep a front with the fixed leading dimension kFrontLd:
    // positions of the own KKT entries in that layout, and the per-front
    // extend-add lists (ExtendList, ldlt_warp.cuh)
    if (Y.update_size >= kExtVecTag || Y.rel_ptr.back() >= kExtVecTag) {
      return fail(S, SLPB_ERR_UNSUPPORTED,
                  "update-matrix storage exceeds 2^30 entries");
    }
    std::vector<int32_t> dst_ld(Y.asm_dst.size());
    std::vector<int32_t> ext_src, ext_dst;
    std::vector<int32_t> ext_begin(Y.n_super, 0), ext_chunks(Y.n_super, 0);
    const int32_t rhs_at = kFrontLd * kFrontCols + kDenseSideDoubles;
    for (int32_t q = 0; q < Y.n_super; ++q) {
      const int32_t F = Y.front_dim[q];
      ext_begin[q] = static_cast<int32_t>(ext_src.size() / 32);
      if (in_top[q]) continue;  // handled by the level kernels
      for (int64_t k = Y.asm_ptr[q]; k < Y.asm_ptr[q + 1]; ++k) {
        dst_ld[k] = Y.asm_dst[k] % F + (Y.asm_dst[k] / F) * kFrontLd;
      }
      for (int64_t ck = Y.child_ptr[q]; ck < Y.child_ptr[q + 1]; ++ck) {
        const int32_t c = Y.child_idx[ck];
        const int32_t mc = Y.front_dim[c] - (Y.super_first[c + 1] - Y.super_first[c]);
        const int32_t* rel = Y.rel_idx.data() + Y.rel_ptr[c];
        for (int32_t j = 0; j < mc; ++j) {
          for (int32_t i = j; i < mc; ++i) {
            ext_src.push_back(static_cast<int32_t>(Y.update_ptr[c] + tri_col(j, mc) + i));
            ext_dst.push_back(rel[i] + rel[j] * kFrontLd);
          }
        }
        for (int32_t i = 0; i < mc; ++i) {
          ext_src.push_back(kExtVecTag | static_cast<int32_t>(Y.rel_ptr[c] + i));
          ext_dst.push_back(rhs_at + rel[i]);
        }
        while (ext_src.size() % 32 != 0) {
          ext_src.push_back(0);
          ext_dst.push_back(-1);
        }
      }
      ext_chunks[q] = static_cast<int32_t>(ext_src.size() / 32) - ext_begin[q];
    }
    if (ext_src.empty()) {
      ext_src.assign(32, 0);
      ext_dst.assign(32, -1);
    }
    CU(S->sy_asm_dst_ld.upload(dst_ld, S->stream));
    CU(S->sy_ext_src.upload(ext_src, S->stream));
    CU(S->sy_ext_dst.upload(ext_dst, S->stream));
    S->ext_begin = std::move(ext_begin);
    S->ext_chunks = std::move(ext_chunks);
  }
  {
    std::vector<int32_t> nchild(Y.n_super);
    for (int32_t q = 0; q < Y.n_super; ++q) {
      nchild[q] = static_cast<int32_t>(Y.child_ptr[q + 1] - Y.child_ptr[q]);
    }
    CU(S->sy_super_parent.upload(Y.super_parent, S->stream));
    CU(S->sy_nchild.upload(nchild, S->stream));
    std::vector<FrontMeta> metas(Y.n_super);
    for (int32_t q = 0; q < Y.n_super; ++q) {
      FrontMeta& fm = metas[q];
      fm.F = Y.front_dim[q];
      fm.c0 = Y.super_first[q];
      fm.np = Y.super_first[q + 1] - Y.super_first[q];
      fm.n_child = nchild[q];
      fm.child_begin = static_cast<int32_t>(Y.child_ptr[q]);
      fm.asm_begin = static_cast<int32_t>(Y.asm_ptr[q]);
      fm.asm_end = static_cast<int32_t>(Y.asm_ptr[q + 1]);
      fm.rel_off = static_cast<int32_t>(Y.rel_ptr[q]);
      fm.panel_off = Y.panel_ptr[q];
      fm.update_off = Y.update_ptr[q];
      fm.rows_off = Y.rows_ptr[q];
      fm.parent = Y.super_parent[q];
      fm.pad = 0;
      fm.ext_begin = S->use_tree ? S->ext_begin[q] : 0;
      fm.ext_chunks = S->use_tree ? S->ext_chunks[q] : 0;
      fm.pad2 = fm.pad3 = 0;
    }
    CU(S->sy_metas.upload(metas, S->stream));
    if (std::getenv("SLPB_TREE_DEBUG")) {
      CU(S->tree_debug.alloc(4 * size_t(Y.n_super)));
      CU(S->tree_debug.zero(S->stream));
    }
    // [0] ticket | 3 × n_super dependency words | [1 + 3 n_super + v] "variant v
    // met an exactly zero pivot" flags of the factor launch
    CU(S->tree_sync.alloc(3 + 3 * size_t(Y.n_super)));
    CU(S->solve_sync.alloc(1 + 3 * size_t(Y.n_super)));
    S->solve_sync_preset = false;
  }
  {
    const int tree_smem =
        kTreeWarps * S->tree_smem_doubles * static_cast<int>(sizeof(double));
    if (S->use_tree) {
      CU(raise_dynamic_smem(k_factor_tree, tree_smem));
      // each warp keeps a 13 KB front workspace: all the shared memory there is
      CU(cudaFuncSetAttribute(k_factor_tree,
                              cudaFuncAttributePreferredSharedMemoryCarveout,
                              cudaSharedmemCarveoutMaxShared));
    }
    // Persistent grids: as many blocks as are resident at once (SMs × blocks
    // per SM at this kernel's register and shared-memory footprint), never more
    // than there is work for. Fronts are handed out by ticket, so blocks of a
    // second wave would only find the tickets gone.
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, S->device);
    int per_sm_factor = 1, per_sm_solve = 1;
    if (S->use_tree) {
      CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
          &per_sm_factor, k_factor_tree, kTreeWarps * 32, tree_smem));
      CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
          &per_sm_solve, k_solve_tree, kTreeWarps * 32, 0));
    }
    const int useful = blocks_for(2 * Y.n_super, kTreeWarps);
    S->tree_blocks =
        std::max(1, std::min(useful, sms * std::max(1, per_sm_factor)));
    // (the solve's warps mostly wait on flags: 16 per SM are plenty)
    S->solve_blocks = std::max(
        1, std::min(useful, sms * std::min(std::max(1, 16 / kTreeWarps),
                                           std::max(1, per_sm_solve))));
  }
  S->tree_sharded = false;
  if (S->world > 1 && S->use_tree && !S->hybrid &&
      !std::getenv("SLPB_NO_TREE_SHARD")) {
    build_tree_shard(Y, S->world, S->tshard);
    const TreeShard& H = S->tshard;
    const int W = S->world;
    std::vector<int32_t> bwd(H.rank_order[S->rank]);
    bwd.insert(bwd.end(), H.top_order.begin(), H.top_order.end());
    CU(S->ts_my_order.upload(H.rank_order[S->rank], S->stream));
    CU(S->ts_top_order.upload(H.top_order, S->stream));
    CU(S->ts_bwd_order.upload(bwd, S->stream));
    CU(S->ts_top_fcount_init.upload(H.top_fcount_init, S->stream));
    // what every rank contributes to the exchange of the subtree roots: the
    // lower triangle of each root's update matrix and its update vector
    std::vector<std::vector<int64_t>> pos(W);
    std::vector<std::vector<int32_t>> sol_idx(W);
    for (int q = 0; q < W; ++q) {
      for (int32_t r : H.rank_roots[q]) {
        const int64_t m =
            Y.front_dim[r] - (Y.super_first[r + 1] - Y.super_first[r]);
        for (int64_t j = 0; j < m; ++j) {
          for (int64_t i = j; i < m; ++i) {
            pos[q].push_back(Y.update_ptr[r] + tri_col(static_cast<int>(j), static_cast<int>(m)) + i);
          }
        }
        for (int64_t i = 0; i < m; ++i) pos[q].push_back((Y.rel_ptr[r] + i) | kPackVecTag);
      }
      for (int32_t f : H.rank_order[q]) {
        for (int32_t c = Y.super_first[f]; c < Y.super_first[f + 1]; ++c) {
          sol_idx[q].push_back(Y.perm[c]);
        }
      }
    }
    S->ts_pack_max = 1;
    S->ts_sol_max = 1;
    for (int q = 0; q < W; ++q) {
      S->ts_pack_max = std::max<int>(S->ts_pack_max, static_cast<int>(pos[q].size()));
      S->ts_sol_max = std::max<int>(S->ts_sol_max, static_cast<int>(sol_idx[q].size()));
    }
    std::vector<int64_t> pos_all(size_t(W) * S->ts_pack_max, 0);
    std::vector<int32_t> sol_all(size_t(W) * S->ts_sol_max, 0);
    S->ts_pack_len_host.assign(W, 0);
    S->ts_sol_len_host.assign(W, 0);
    for (int q = 0; q < W; ++q) {
      std::copy(pos[q].begin(), pos[q].end(), pos_all.begin() + size_t(q) * S->ts_pack_max);
      std::copy(sol_idx[q].begin(), sol_idx[q].end(), sol_all.begin() + size_t(q) * S->ts_sol_max);
      S->ts_pack_len_host[q] = static_cast<int32_t>(pos[q].size());
      S->ts_sol_len_host[q] = static_cast<int32_t>(sol_idx[q].size());
    }
    CU(S->ts_pack_pos.upload(pos_all, S->stream));
    CU(S->ts_sol_idx.upload(sol_all, S->stream));
    CU(S->ts_pack_len.upload(S->ts_pack_len_host, S->stream));
    CU(S->ts_sol_len.upload(S->ts_sol_len_host, S->stream));
    CU(S->ts_pack_buf.alloc(size_t(W) * 2 * (kPackStats + S->ts_pack_max)));
    CU(S->ts_pack_buf.zero(S->stream));
    CU(S->ts_sol_buf.alloc(size_t(W) * S->ts_sol_max));
    CU(S->ts_sol_buf.zero(S->stream));
    for (auto& e : S->cev) {
      if (!e) CU(cudaEventCreate(&e));
    }
    S->tree_sharded = true;
  }
  CU(cudaStreamSynchronize(S->stream));
  SymbolicView& V = S->sview;
  V.dim = Y.dim;
  V.n_super = Y.n_super;
  V.super_first = S->sy_super_first.p;
  V.front_dim = S->sy_front_dim.p;
  V.rows_ptr = S->sy_rows_ptr.p;
  V.rows_idx = S->sy_rows_idx.p;
  V.panel_ptr = S->sy_panel_ptr.p;
  V.update_ptr = S->sy_update_ptr.p;
  V.child_ptr = S->sy_child_ptr.p;
  V.child_idx = S->sy_child_idx.p;
  V.rel_ptr = S->sy_rel_ptr.p;
  V.rel_idx = S->sy_rel_idx.p;
  V.asm_ptr = S->sy_asm_ptr.p;
  V.asm_src = S->sy_asm_src.p;
  V.asm_dst = S->sy_asm_dst.p;
  V.col_is_primal = S->sy_col_is_primal.p;
  V.perm = S->sy_perm.p;
  S->analyzed = true;
  if (stats) {
    stats->dim = Y.dim;
    stats->nnz_kkt = S->recipe.K.nnz();
    stats->nnz_l = Y.nnz_l;
    stats->nnz_l_stored = Y.panel_size;
    stats->n_supernodes = Y.n_super;
    stats->n_levels = Y.n_levels;
    stats->max_front = Y.max_front;
    stats->etree_height = Y.etree_height;
  }
  return SLPB_OK;
}

int slpb_get_permutation(const slpb_solver* S, int32_t* perm) {
  if (!S || !S->analyzed || !perm) return SLPB_ERR_STATE;
  std::memcpy(perm, S->sym.perm.data(), S->sym.dim * sizeof(int32_t));
  return SLPB_OK;
}

int slpb_set_iterate(slpb_solver* S, const double* x, const double* sl,
                     const double* y, const double* z) {
  if (!S || !S->finalized) return SLPB_ERR_STATE;
  CU(cudaSetDevice(S->device));
  S->rhs_ready = false;
  auto put = [&](DevBuf<double>& b, const double* src) -> cudaError_t {
    if (b.n == 0) return cudaSuccess;
    if (!src) return cudaErrorInvalidValue;
    S->counters.h2d_bytes += b.n * sizeof(double);
    return cudaMemcpyAsync(b.p, src, b.n * sizeof(double),
                           cudaMemcpyHostToDevice, S->stream);
  };
  CU(put(S->x, x));
  CU(put(S->s, sl));
  CU(put(S->y, y));
  CU(put(S->z, z));
  CU(cudaStreamSynchronize(S->stream));
  return SLPB_OK;
}

int slpb_get_iterate(slpb_solver* S, double* x, double* sl, double* y,
                     double* z) {
  if (!S || !S->finalized) return SLPB_ERR_STATE;
  CU(cudaSetDevice(S->device));
  auto get = [&](DevBuf<double>& b, double* dst) -> cudaError_t {
    if (b.n == 0 || !dst) return cudaSuccess;
    S->counters.d2h_bytes += b.n * sizeof(double);
    return cudaMemcpyAsync(dst, b.p, b.n * sizeof(double),
                           cudaMemcpyDeviceToHost, S->stream);
  };
  CU(get(S->x, x));
  CU(get(S->s, sl));
  CU(get(S->y, y));
  CU(get(S->z, z));
  CU(cudaStreamSynchronize(S->stream));
  return SLPB_OK;
}

int slpb_eval_current(slpb_solver* S, int derivatives,
                      slpb_point_info* info) {
  if (!S || !S->finalized || !info) return SLPB_ERR_STATE;
  CU(cudaSetDevice(S->device));
  S->rhs_ready = false;
  int rc;
  if ((rc = refresh_leaves(S, S->x.p, S->y.p, S->z.p, S->leaf_cur.p))) return rc;
  if (derivatives != 2) {
    if ((rc = eval_values(S, S->leaf_cur.p, S->vals_cur.p))) return rc;
  }
  if ((rc = point_info(S, S->vals_cur.p, S->s.p, S->d_results.p))) return rc;
  int nres = 6;
  if (derivatives != 0) {
    if ((rc = eval_derivs(S, S->leaf_cur.p))) return rc;
    k_deriv_finite<<<red_blocks(S->ad.off_h + S->ad.H.nnz()), kReduceThreads,
                     0, S->stream>>>(
        S->dvals.p, S->ad.off_ae, S->ad.off_ai, S->ad.off_h,
        S->ad.off_h + S->ad.H.nnz(), red_buf(S), S->d_results.p + 6);
    ++S->counters.kernel_launches;
    nres = 7;
  }
  if ((rc = fetch_results(S, nres))) return rc;
  fill_point_info(S->h_results, info);
  if (derivatives != 0) info->finite |= static_cast<int32_t>(S->h_results[6]);
  return SLPB_OK;
}

int slpb_kkt_stats_current(slpb_solver* S, double mu, slpb_kkt_stats* out) {
  if (!S || !S->finalized || !out) return SLPB_ERR_STATE;
  CU(cudaSetDevice(S->device));
  return kkt_stats(S, S->vals_cur.p + 1, S->vals_cur.p + 1 + S->me, S->x.p,
                   S->s.p, S->y.p, S->z.p, mu, out);
}

int slpb_kkt_stats_trial(slpb_solver* S, double mu, slpb_kkt_stats* out) {
  if (!S || !S->finalized || !out) return SLPB_ERR_STATE;
  CU(cudaSetDevice(S->device));
  // g, A_e, A_i at the trial point (interior_point.hpp:704-706). This
  // overwrites the derivative arrays; the driver re-linearises afterwards.
  S->rhs_ready = false;
  int rc;
  if ((rc = refresh_leaves(S, S->tx.p, S->ty.p, S->tz.p, S->leaf_trial.p))) {
    return rc;
  }
  if ((rc = eval_derivs(S, S->leaf_trial.p))) return rc;
  return kkt_stats(S, S->vals_trial.p + 1, S->vals_trial.p + 1 + S->me,
                   S->tx.p, S->ts.p, S->ty.p, S->tz.p, mu, out);
}

namespace {

/// Assembles (optionally) and factors one or two regularisations of the KKT
/// matrix in one dependency-driven launch.
int factor_impl(slpb_solver* S, int n_variants, const double* delta,
                const double* gamma, int reassemble, slpb_factor_info* info) {
  CU(cudaSetDevice(S->device));
  const Symbolic& Y = S->sym;
  // The first attempt of an iteration assembles the lhs and prepares the launch
  // (statistics, ticket, dependency counters, dead flags) in ONE kernel; Σ is
  // already there when slpb_prepare_rhs has just run for this iterate.
  const bool fuse_init = reassemble && S->group == nullptr;
  // a fused solve is expected behind this factorisation: its dependency words
  // are preset here, so that the solve is ONE launch
  const bool preset_solve = S->use_tree && !S->tree_sharded && !S->hybrid &&
                            S->rhs_ready && S->group == nullptr;
  int32_t* preset_ptr = preset_solve ? S->solve_sync.p : nullptr;
  const int init_words =
      S->use_tree ? std::max(1 + 2 * Y.n_super, preset_solve ? 1 + 3 * Y.n_super : 0)
                  : 16;
  if (reassemble) {
    CU(timer_begin(S, 2));
    if (S->mi > 0 && !S->rhs_ready) {
      k_sigma_t<<<blocks_for(S->mi, 256), 256, 0, S->stream>>>(
          S->s.p, S->z.p, S->vals_cur.p + 1 + S->me, nullptr, 0.0, 0, S->mi,
          S->sinv.p, S->sigma.p, S->tvec.p);
      ++S->counters.kernel_launches;
    }
    const int nnz = static_cast<int>(S->recipe.K.nnz());
    if (fuse_init) {
      k_kkt_assemble_init<<<blocks_for(std::max(nnz, init_words), 256), 256, 0,
                            S->stream>>>(
          S->k_h_idx.p, S->k_ae_idx.p, S->k_prod_ptr.p, S->k_prod_a.p,
          S->k_prod_b.p, S->k_prod_row.p, S->dvals.p + S->ad.off_h,
          S->dvals.p + S->ad.off_ae, S->dvals.p + S->ad.off_ai, S->sigma.p, nnz,
          S->Kval.p, S->fstats.p, S->tree_sync.p, Y.n_super, S->use_tree ? 1 : 0,
          preset_ptr);
    } else {
      k_kkt_assemble<<<blocks_for(nnz, 256), 256, 0, S->stream>>>(
          S->k_h_idx.p, S->k_ae_idx.p, S->k_prod_ptr.p, S->k_prod_a.p,
          S->k_prod_b.p, S->k_prod_row.p, S->dvals.p + S->ad.off_h,
          S->dvals.p + S->ad.off_ae, S->dvals.p + S->ad.off_ai, S->sigma.p, nnz,
          S->Kval.p);
    }
    ++S->counters.kernel_launches;
    CU(timer_end(S, 2));
  }
  if (S->group != nullptr) {
    // member of a batching group: the lhs values are ready on this stream;
    // the factorisation itself runs in the group's next batched launch
    CU(cudaStreamSynchronize(S->stream));
    harvest_timers(S);
    return group_factor(S, n_variants, delta, gamma, info);
  }
  CU(timer_begin(S, 3));
  // stats per variant: n_pos n_neg n_zero zero_pivot | min|D| bits (+inf);
  // ticket, dependency counters and dead flags of the tree kernel
  if (!fuse_init) {
    k_init_factor<<<blocks_for(init_words, 256), 256, 0, S->stream>>>(
        S->fstats.p, S->tree_sync.p, Y.n_super, S->use_tree ? 1 : 0, preset_ptr);
    ++S->counters.kernel_launches;
  }
  S->solve_sync_preset = preset_solve;
  if (S->use_tree) {
    const TreeView T = tree_view(S);
    const int smem =
        kTreeWarps * S->tree_smem_doubles * static_cast<int>(sizeof(double));
    const FactorPair pair{n_variants,
                          n_variants == 2 ? delta[1] : 0.0,
                          n_variants == 2 ? gamma[1] : 0.0,
                          Y.panel_size,
                          Y.update_size,
                          static_cast<int64_t>(Y.rel_ptr.back()),
                          Y.dim,
                          // (hybrid: the level kernels of the top do not carry
                          // a right-hand side; the solve runs both halves)
                          S->rhs_ready && !S->hybrid ? S->rhs.p : nullptr,
                          S->xperm.p,
                          S->uvecs.p,
                          S->factor_arith};
    for (int v = 0; v < 2; ++v) {
      S->fwd_valid[v] = S->rhs_ready && !S->hybrid && v < n_variants;
    }
    if (S->hybrid) {
      // everything below the oversized fronts in one tree launch, then the
      // top level by level, one block per front
      TreeView Ts = T;
      Ts.order = S->hy_small_order.p;
      Ts.n_order = S->n_small;
      k_factor_tree<<<S->tree_blocks, kTreeWarps * 32, smem, S->stream>>>(
          Ts, S->Kval.p, delta[0], gamma[0], pair, S->panels.p, S->updates.p,
          S->D.p, S->fstats.p);
      const int top_smem = S->front_smem_bytes;
      for (int v = 0; v < n_variants; ++v) {
        for (size_t L = 0; L + 1 < S->hy_level_off.size(); ++L) {
          const int cnt = S->hy_level_off[L + 1] - S->hy_level_off[L];
          k_factor_level<<<cnt, kFrontThreads, top_smem, S->stream>>>(
              S->sview, S->hy_top_supers.p + S->hy_level_off[L], S->Kval.p,
              delta[v], gamma[v], S->panels.p + v * Y.panel_size,
              S->updates.p + v * Y.update_size, S->D.p + size_t(v) * Y.dim,
              S->fstats.p + 8 * v, S->factor_arith, S->front_scratch.p,
              S->front_slot.p, S->front_stride, S->front_cap_doubles);
          ++S->counters.kernel_launches;
        }
      }
    } else if (!S->tree_sharded) {
      k_factor_tree<<<S->tree_blocks, kTreeWarps * 32, smem, S->stream>>>(
          T, S->Kval.p, delta[0], gamma[0], pair, S->panels.p, S->updates.p,
          S->D.p, S->fstats.p);
    } else {
      // Sharded: own subtrees → one small all-gather of the
      // subtree roots (update matrices, update vectors, inertia counts) → the
      // replicated top, identically on every rank.
      const TreeShard& H = S->tshard;
      const int n_mine = static_cast<int>(H.rank_order[S->rank].size());
      const int n_top = static_cast<int>(H.top_order.size());
      TreeView Tm = T;
      Tm.order = S->ts_my_order.p;
      Tm.n_order = n_mine;
      if (n_mine > 0) {
        k_factor_tree<<<S->tree_blocks, kTreeWarps * 32, smem, S->stream>>>(
            Tm, S->Kval.p, delta[0], gamma[0], pair, S->panels.p, S->updates.p,
            S->D.p, S->fstats.p);
        ++S->counters.kernel_launches;
      }
      int rc = exchange_roots(S, n_variants, true);
      if (rc) return rc;
      CU(cudaMemsetAsync(S->tree_sync.p, 0, 4, S->stream));  // ticket
      for (int v = 0; v < n_variants; ++v) {
        CU(cudaMemcpyAsync(S->tree_sync.p + 1 + size_t(v) * Y.n_super,
                           S->ts_top_fcount_init.p, size_t(Y.n_super) * 4,
                           cudaMemcpyDeviceToDevice, S->stream));
      }
      TreeView Tt = T;
      Tt.order = S->ts_top_order.p;
      Tt.n_order = n_top;
      const int blocks = std::max(
          1, std::min(S->tree_blocks,
                      blocks_for(int64_t(n_variants) * n_top, kTreeWarps)));
      k_factor_tree<<<blocks, kTreeWarps * 32, smem, S->stream>>>(
          Tt, S->Kval.p, delta[0], gamma[0], pair, S->panels.p, S->updates.p,
          S->D.p, S->fstats.p);
    }
    ++S->counters.kernel_launches;
  } else {
    S->fwd_valid[0] = S->fwd_valid[1] = false;
    const int smem = S->front_smem_bytes;
    for (int v = 0; v < n_variants; ++v) {
      for (int L = 0; L < Y.n_levels; ++L) {
        const int cnt = Y.level_ptr[L + 1] - Y.level_ptr[L];
        k_factor_level<<<cnt, kFrontThreads, smem, S->stream>>>(
            S->sview, S->sy_level_supers.p + Y.level_ptr[L], S->Kval.p,
            delta[v], gamma[v], S->panels.p + v * Y.panel_size,
            S->updates.p + v * Y.update_size, S->D.p + size_t(v) * Y.dim,
            S->fstats.p + 8 * v, S->factor_arith, S->front_scratch.p,
            S->front_slot.p, S->front_stride, S->front_cap_doubles);
      }
      S->counters.kernel_launches += Y.n_levels;
    }
  }
  CU(timer_end(S, 3));
  CU(cudaGetLastError());
  // (into the pinned result buffer: a pageable destination would be staged)
  int32_t* host_stats = reinterpret_cast<int32_t*>(S->h_results);
  static_assert(kResultDoubles * sizeof(double) >= 16 * sizeof(int32_t));
  CU(cudaMemcpyAsync(host_stats, S->fstats.p, 16 * sizeof(int32_t),
                     cudaMemcpyDeviceToHost, S->stream));
  CU(cudaStreamSynchronize(S->stream));
  harvest_timers(S);
  S->counters.d2h_bytes += 16 * sizeof(int32_t);
  for (int v = 0; v < n_variants; ++v) {
    info[v].n_pos = host_stats[8 * v + 0];
    info[v].n_neg = host_stats[8 * v + 1];
    info[v].n_zero = host_stats[8 * v + 2];
    info[v].zero_pivot = host_stats[8 * v + 3];
    std::memcpy(&info[v].min_abs_d, &host_stats[8 * v + 4], 8);
    if (info[v].zero_pivot) {
      S->fwd_valid[v] = false;  // factor was abandoned
    } else {
      ++S->counters.factorizations_completed;
    }
  }
  S->counters.factorizations += n_variants;
  S->factor_sel = 0;
  return SLPB_OK;
}

}  // namespace

int slpb_factor(slpb_solver* S, double delta, double gamma, int reassemble,
                slpb_factor_info* info) {
  if (!S || !S->analyzed || !info) return SLPB_ERR_STATE;
  return factor_impl(S, 1, &delta, &gamma, reassemble, info);
}

int slpb_factor_pair(slpb_solver* S, const double delta[2],
                     const double gamma[2], int reassemble,
                     slpb_factor_info info[2]) {
  if (!S || !S->analyzed || !info || !delta || !gamma) return SLPB_ERR_STATE;
  return factor_impl(S, 2, delta, gamma, reassemble, info);
}

int slpb_select_factor(slpb_solver* S, int which) {
  if (!S || !S->analyzed) return SLPB_ERR_STATE;
  if (which != 0 && which != 1) return SLPB_ERR_ARGUMENT;
  S->factor_sel = which;
  return SLPB_OK;
}

int slpb_prepare_rhs(slpb_solver* S, double mu) {
  if (!S || !S->analyzed) return SLPB_ERR_STATE;
  CU(cudaSetDevice(S->device));
  int rc = build_rhs(S, mu, nullptr, S->vals_cur.p + 1);
  if (rc) return rc;
  S->rhs_ready = true;
  S->rhs_mu = mu;
  S->fwd_valid[0] = S->fwd_valid[1] = false;
  return SLPB_OK;
}

int slpb_solve(slpb_solver* S, double mu, double tau, slpb_step_info* info) {
  if (!S || !S->analyzed || !info) return SLPB_ERR_STATE;
  CU(cudaSetDevice(S->device));
  return solve_into(S, mu, tau, false, nullptr, S->vals_cur.p + 1, S->px.p,
                    S->ps.p, S->py.p, S->pz.p, info);
}

int slpb_soc_begin(slpb_solver* S) {
  if (!S || !S->analyzed) return SLPB_ERR_STATE;
  CU(cudaSetDevice(S->device));
  const int m = std::max(S->me, S->mi);
  if (m > 0) {
    k_soc_begin<<<blocks_for(m, 256), 256, 0, S->stream>>>(
        S->vals_cur.p + 1, S->vals_cur.p + 1 + S->me, S->s.p, S->me, S->mi,
        S->ce_soc.p, S->cis_soc.p);
    ++S->counters.kernel_launches;
  }
  CU(cudaGetLastError());
  return SLPB_OK;
}

int slpb_soc_iterate(slpb_solver* S, double mu, double tau, double alpha_soc,
                     slpb_step_info* info) {
  if (!S || !S->analyzed || !info) return SLPB_ERR_STATE;
  CU(cudaSetDevice(S->device));
  const int m = std::max(S->me, S->mi);
  if (m > 0) {
    k_soc_accumulate<<<blocks_for(m, 256), 256, 0, S->stream>>>(
        S->vals_trial.p + 1, S->vals_trial.p + 1 + S->me, S->ts.p, alpha_soc,
        S->me, S->mi, S->ce_soc.p, S->cis_soc.p);
    ++S->counters.kernel_launches;
  }
  return solve_into(S, mu, tau, true, S->cis_soc.p, S->ce_soc.p, S->spx.p,
                    S->sps.p, S->spy.p, S->spz.p, info);
}

namespace {
/// Trial point + f, c_e, c_i there + its reductions at d_results[offset..];
/// no host round trip.
int trial_enqueue(slpb_solver* S, double alpha, double alpha_z,
                  const double* alpha_dev, int dual_uses_primal,
                  int which_step, int slack_from_ci, int offset) {
  const int n = S->n, me = S->me, mi = S->mi;
  const bool soc = which_step != 0;
  const int m = std::max(n, std::max(me, mi));
  k_trial_point<<<blocks_for(m, 256), 256, 0, S->stream>>>(
      S->x.p, S->s.p, S->y.p, S->z.p, soc ? S->spx.p : S->px.p,
      soc ? S->sps.p : S->ps.p, soc ? S->spy.p : S->py.p,
      soc ? S->spz.p : S->pz.p, alpha, alpha_z, alpha_dev, dual_uses_primal, n,
      me, mi, S->tx.p, S->ts.p, S->ty.p, S->tz.p, S->leaf_trial.p);
  ++S->counters.kernel_launches;
  int rc;
  if ((rc = eval_values(S, S->leaf_trial.p, S->vals_trial.p))) return rc;
  if (slack_from_ci && mi > 0) {
    k_copy<<<blocks_for(mi, 256), 256, 0, S->stream>>>(
        S->vals_trial.p + 1 + me, S->ts.p, mi);
    ++S->counters.kernel_launches;
  }
  return point_info(S, S->vals_trial.p, S->ts.p, S->d_results.p + offset);
}
}  // namespace

int slpb_trial(slpb_solver* S, double alpha, double alpha_z, int which_step,
               int slack_from_ci, slpb_point_info* info) {
  if (!S || !S->analyzed || !info) return SLPB_ERR_STATE;
  CU(cudaSetDevice(S->device));
  int rc;
  if ((rc = trial_enqueue(S, alpha, alpha_z, nullptr, 0, which_step,
                                slack_from_ci, 0))) {
    return rc;
  }
  if ((rc = fetch_results(S, 6))) return rc;
  fill_point_info(S->h_results, info);
  return SLPB_OK;
}

int slpb_solve_trial(slpb_solver* S, double mu, double tau,
                     int dual_uses_primal_alpha, int slack_from_ci,
                     slpb_step_info* step, slpb_point_info* trial) {
  if (!S || !S->analyzed || !step || !trial) return SLPB_ERR_STATE;
  CU(cudaSetDevice(S->device));
  int rc;
  if ((rc = solve_into(S, mu, tau, false, nullptr, S->vals_cur.p + 1, S->px.p,
                       S->ps.p, S->py.p, S->pz.p, nullptr))) {
    return rc;
  }
  if ((rc = trial_enqueue(S, 0.0, 0.0, S->d_results.p + kResStep,
                                dual_uses_primal_alpha, 0, slack_from_ci,
                                kResPoint))) {
    return rc;
  }
  if ((rc = fetch_results(S, kResPoint + 6))) return rc;
  fill_step_info(S->h_results + kResStep, step);
  fill_point_info(S->h_results + kResPoint, trial);
  return SLPB_OK;
}

int slpb_probe_point(slpb_solver* S, const double* x, const double* sl,
                     slpb_point_info* info) {
  if (!S || !S->finalized || !info || !x) return SLPB_ERR_STATE;
  CU(cudaSetDevice(S->device));
  const int n = S->n, me = S->me, mi = S->mi;
  if (mi > 0 && !sl) return SLPB_ERR_ARGUMENT;
  CU(cudaMemcpyAsync(S->tx.p, x, n * sizeof(double), cudaMemcpyHostToDevice,
                     S->stream));
  if (mi > 0) {
    CU(cudaMemcpyAsync(S->ts.p, sl, mi * sizeof(double),
                       cudaMemcpyHostToDevice, S->stream));
    CU(cudaMemcpyAsync(S->tz.p, S->z.p, mi * sizeof(double),
                       cudaMemcpyDeviceToDevice, S->stream));
  }
  if (me > 0) {
    CU(cudaMemcpyAsync(S->ty.p, S->y.p, me * sizeof(double),
                       cudaMemcpyDeviceToDevice, S->stream));
  }
  S->counters.h2d_bytes += (n + mi) * sizeof(double);
  int rc;
  if ((rc = refresh_leaves(S, S->tx.p, S->ty.p, S->tz.p, S->leaf_trial.p))) {
    return rc;
  }
  if ((rc = eval_values(S, S->leaf_trial.p, S->vals_trial.p))) return rc;
  if ((rc = point_info(S, S->vals_trial.p, S->ts.p, S->d_results.p))) return rc;
  if ((rc = fetch_results(S, 6))) return rc;
  fill_point_info(S->h_results, info);
  return SLPB_OK;
}

int slpb_multiplier_estimate(slpb_solver* S, double mu, slpb_factor_info* info) {
  if (!S || !S->analyzed || !info) return SLPB_ERR_STATE;
  CU(cudaSetDevice(S->device));
  const int n = S->n, me = S->me, mi = S->mi;
  S->rhs_ready = false;
  if (mi > 0) {
    k_inv_square<<<blocks_for(mi, 256), 256, 0, S->stream>>>(S->s.p, mi,
                                                            S->sigma.p);
    ++S->counters.kernel_launches;
  }
  const int nnz = static_cast<int>(S->recipe.K.nnz());
  k_kkt_assemble_identity<<<blocks_for(nnz, 256), 256, 0, S->stream>>>(
      S->k_diag_flag.p, S->k_ae_idx.p, S->k_prod_ptr.p, S->k_prod_a.p,
      S->k_prod_b.p, S->k_prod_row.p, S->dvals.p + S->ad.off_ae,
      S->dvals.p + S->ad.off_ai, S->sigma.p, nnz, S->Kval.p);
  k_rhs_estimate<<<blocks_for(S->dim, 256), 256, 0, S->stream>>>(
      ai_view(S), S->dvals.p + S->ad.off_g, S->s.p, mu, n, me, S->rhs.p);
  S->counters.kernel_launches += 2;
  CU(cudaGetLastError());
  // the (2,2) block is zero: a tiny γ keeps multipliers of degree-1 rows (which
  // the ordering may eliminate first) from meeting an exactly zero pivot
  const double delta = 0.0, gamma = 1e-10;
  int rc = factor_impl(S, 1, &delta, &gamma, /*reassemble=*/0, info);
  if (rc) return rc;
  S->fwd_valid[0] = S->fwd_valid[1] = false;
  if ((rc = launch_solve(S, false))) return rc;
  const int m = std::max(me, mi);
  if (m > 0) {
    k_estimate_finish<<<blocks_for(m, 256), 256, 0, S->stream>>>(
        S->sol.p, S->ai_rowptr.p, S->ai_rcol.p, S->ai_ridx.p,
        S->dvals.p + S->ad.off_ai, S->s.p, mu, n, me, mi, S->y.p, S->z.p);
    ++S->counters.kernel_launches;
  }
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(S->stream));
  harvest_timers(S);
  return SLPB_OK;
}

int slpb_accept(slpb_solver* S, double mu) {
  if (!S || !S->analyzed) return SLPB_ERR_STATE;
  S->rhs_ready = false;
  CU(cudaSetDevice(S->device));
  auto cp = [&](DevBuf<double>& dst, DevBuf<double>& src) -> cudaError_t {
    if (dst.n == 0) return cudaSuccess;
    return cudaMemcpyAsync(dst.p, src.p, dst.n * sizeof(double),
                           cudaMemcpyDeviceToDevice, S->stream);
  };
  CU(cp(S->x, S->tx));
  CU(cp(S->s, S->ts));
  CU(cp(S->y, S->ty));
  CU(cp(S->z, S->tz));
  CU(cp(S->vals_cur, S->vals_trial));
  if (S->mi > 0) {
    k_clamp_z<<<blocks_for(S->mi, 256), 256, 0, S->stream>>>(S->s.p, mu, S->mi,
                                                            S->z.p);
    ++S->counters.kernel_launches;
  }
  CU(cudaGetLastError());
  return SLPB_OK;
}

int slpb_accept_relinearize(slpb_solver* S, double mu, int32_t* finite,
                            slpb_kkt_stats* stats) {
  if (!S || !S->analyzed || !finite || !stats) return SLPB_ERR_STATE;
  int rc;
  {
    // slpb_accept + the leaves of the accepted point, in one launch
    S->rhs_ready = false;
    CU(cudaSetDevice(S->device));
    const int n = S->n, me = S->me, mi = S->mi;
    const int m = std::max(std::max(n, 1 + me + mi), std::max(me, mi));
    k_accept<<<blocks_for(m, 256), 256, 0, S->stream>>>(
        S->tx.p, S->ts.p, S->ty.p, S->tz.p, S->vals_trial.p, S->d_c.p, mu, n, me,
        mi, S->x.p, S->s.p, S->y.p, S->z.p, S->vals_cur.p, S->leaf_cur.p);
    ++S->counters.kernel_launches;
    CU(cudaGetLastError());
  }
  if ((rc = eval_derivs(S, S->leaf_cur.p))) return rc;
  // error reductions and the finiteness scan of the new derivatives, one launch
  if ((rc = kkt_stats_enqueue(S, S->vals_cur.p + 1, S->vals_cur.p + 1 + S->me,
                              S->x.p, S->s.p, S->y.p, S->z.p, mu, kResKkt,
                              /*with_deriv_finite=*/true))) {
    return rc;
  }
  if ((rc = fetch_results(S, kResKkt + 25))) return rc;
  *finite = static_cast<int32_t>(S->h_results[kResPoint]);
  fill_kkt_stats(S->h_results + kResKkt, stats);
  return SLPB_OK;
}

int slpb_array_size(const slpb_solver* S, int which, int64_t* count) {
  if (!S || !S->finalized || !count) return SLPB_ERR_STATE;
  switch (which) {
    case SLPB_ARR_X: case SLPB_ARR_G: case SLPB_ARR_P_X: case SLPB_ARR_TRIAL_X:
      *count = S->n;
      break;
    case SLPB_ARR_S: case SLPB_ARR_Z: case SLPB_ARR_C_I: case SLPB_ARR_P_S:
    case SLPB_ARR_P_Z: case SLPB_ARR_TRIAL_S: case SLPB_ARR_TRIAL_Z:
    case SLPB_ARR_TRIAL_C_I:
      *count = S->mi;
      break;
    case SLPB_ARR_Y: case SLPB_ARR_C_E: case SLPB_ARR_P_Y:
    case SLPB_ARR_TRIAL_Y: case SLPB_ARR_TRIAL_C_E:
      *count = S->me;
      break;
    case SLPB_ARR_A_E_VAL: *count = S->ad.A_e.nnz(); break;
    case SLPB_ARR_A_I_VAL: *count = S->ad.A_i.nnz(); break;
    case SLPB_ARR_H_VAL: *count = S->ad.H.nnz(); break;
    case SLPB_ARR_KKT_VAL: *count = S->recipe.K.nnz(); break;
    case SLPB_ARR_D: case SLPB_ARR_RHS: *count = S->dim; break;
    default: return SLPB_ERR_ARGUMENT;
  }
  return SLPB_OK;
}

int slpb_download(slpb_solver* S, int which, double* dst) {
  if (!S || !S->finalized || !dst) return SLPB_ERR_STATE;
  CU(cudaSetDevice(S->device));
  int64_t count = 0;
  int rc = slpb_array_size(S, which, &count);
  if (rc) return rc;
  const double* src = nullptr;
  const int me = S->me;
  switch (which) {
    case SLPB_ARR_X: src = S->x.p; break;
    case SLPB_ARR_S: src = S->s.p; break;
    case SLPB_ARR_Y: src = S->y.p; break;
    case SLPB_ARR_Z: src = S->z.p; break;
    case SLPB_ARR_G: src = S->dvals.p + S->ad.off_g; break;
    case SLPB_ARR_C_E: src = S->vals_cur.p + 1; break;
    case SLPB_ARR_C_I: src = S->vals_cur.p + 1 + me; break;
    case SLPB_ARR_A_E_VAL: src = S->dvals.p + S->ad.off_ae; break;
    case SLPB_ARR_A_I_VAL: src = S->dvals.p + S->ad.off_ai; break;
    case SLPB_ARR_H_VAL: src = S->dvals.p + S->ad.off_h; break;
    case SLPB_ARR_KKT_VAL: src = S->Kval.p; break;
    case SLPB_ARR_D:
      if (S->group != nullptr) return group_download_d(S, dst);
      src = S->D.p + size_t(S->factor_sel) * S->dim;
      break;
    case SLPB_ARR_RHS: src = S->rhs.p; break;
    case SLPB_ARR_P_X: src = S->px.p; break;
    case SLPB_ARR_P_S: src = S->ps.p; break;
    case SLPB_ARR_P_Y: src = S->py.p; break;
    case SLPB_ARR_P_Z: src = S->pz.p; break;
    case SLPB_ARR_TRIAL_X: src = S->tx.p; break;
    case SLPB_ARR_TRIAL_S: src = S->ts.p; break;
    case SLPB_ARR_TRIAL_Y: src = S->ty.p; break;
    case SLPB_ARR_TRIAL_Z: src = S->tz.p; break;
    case SLPB_ARR_TRIAL_C_E: src = S->vals_trial.p + 1; break;
    case SLPB_ARR_TRIAL_C_I: src = S->vals_trial.p + 1 + me; break;
    default: return SLPB_ERR_ARGUMENT;
  }
  if (count > 0) {
    if (!src) return fail(S, SLPB_ERR_STATE, "array not available yet");
    CU(cudaMemcpyAsync(dst, src, count * sizeof(double),
                       cudaMemcpyDeviceToHost, S->stream));
    CU(cudaStreamSynchronize(S->stream));
    S->counters.d2h_bytes += count * sizeof(double);
  }
  return SLPB_OK;
}

int slpb_pattern(const slpb_solver* S, int which, int32_t* rows, int32_t* cols,
                 int64_t* nnz, int32_t* colptr, int32_t* rowidx) {
  if (!S || !S->finalized) return SLPB_ERR_STATE;
  const Pattern* p = nullptr;
  switch (which) {
    case SLPB_OUT_A_E: p = &S->ad.A_e; break;
    case SLPB_OUT_A_I: p = &S->ad.A_i; break;
    case SLPB_OUT_H_C: case SLPB_OUT_H_F: p = &S->ad.H; break;
    case -1: p = &S->recipe.K; break;
    default: return SLPB_ERR_ARGUMENT;
  }
  if (rows) *rows = p->rows;
  if (cols) *cols = p->cols;
  if (nnz) *nnz = p->nnz();
  if (colptr) {
    std::memcpy(colptr, p->colptr.data(), p->colptr.size() * sizeof(int32_t));
  }
  if (rowidx && p->nnz() > 0) {
    std::memcpy(rowidx, p->rowidx.data(), p->nnz() * sizeof(int32_t));
  }
  return SLPB_OK;
}

/* Undocumented debugging aid (a build with -DSLPB_SWEEP_STAMPS): clock64() at
 * the phase boundaries of blocks 0 and 75 (out[64..]) of the last 512-thread
 * k_ad_sweep launch — with 157 tasks on 148 SMs block 0 shares its SM:
 * [0] entry, [1] tables in shared memory, [2] leaves loaded, [3 + b] after the
 * barrier of instruction block b, [62] outputs stored. Not part of
 * include/slpb.h. */
int slpb_debug_sweep_stamps(long long* out) {
#ifdef SLPB_SWEEP_STAMPS
  if (cudaDeviceSynchronize() != cudaSuccess) return SLPB_ERR_CUDA;
  void* addr = nullptr;
  if (cudaGetSymbolAddress(&addr, g_sweep_stamps) != cudaSuccess) return SLPB_ERR_CUDA;
  return cudaMemcpy(out, addr, sizeof(long long) * 128, cudaMemcpyDeviceToHost) == cudaSuccess
             ? SLPB_OK
             : SLPB_ERR_CUDA;
#else
  (void)out;
  return SLPB_ERR_UNSUPPORTED;
#endif
}

/* Undocumented debugging aid (SLPB_TREE_DEBUG=1): per front, globaltimer ns at
 * ticket, after the children wait, and at completion of the last
 * factorisation; level of each front. Not part of include/slpb.h. */
int slpb_debug_tree(slpb_solver* S, unsigned long long* stamps, int32_t* level,
                    int32_t* parent) {
  if (!S || !S->analyzed || !S->tree_debug.p) return SLPB_ERR_STATE;
  CU(cudaSetDevice(S->device));
  CU(cudaStreamSynchronize(S->stream));
  CU(cudaMemcpy(stamps, S->tree_debug.p, S->tree_debug.n * 8,
                cudaMemcpyDeviceToHost));
  std::memcpy(level, S->sym.super_level.data(), S->sym.n_super * 4);
  std::memcpy(parent, S->sym.super_parent.data(), S->sym.n_super * 4);
  return SLPB_OK;
}

int slpb_get_counters(const slpb_solver* S, slpb_counters* out) {
  if (!S || !out) return SLPB_ERR_ARGUMENT;
  *out = S->counters;
  return SLPB_OK;
}

int slpb_last_device_ms(slpb_solver* S, int which, float* ms) {
  if (!S || !ms || which < 0 || which > 4) return SLPB_ERR_ARGUMENT;
  CU(cudaSetDevice(S->device));
  CU(cudaStreamSynchronize(S->stream));
  harvest_timers(S);
  *ms = S->last_ms[which];
  return SLPB_OK;
}

int slpb_get_comm_stats(slpb_solver* S, slpb_comm_stats* out) {
  if (!S || !out) return SLPB_ERR_ARGUMENT;
  CU(cudaSetDevice(S->device));
  CU(cudaStreamSynchronize(S->stream));
  harvest_timers(S);
  *out = S->comm_stats;
  // (not every call is timed: scale the timed ones to the call count)
  for (int w = 0; w < 3; ++w) {
    if (S->comm_timed[w] > 0) {
      out->total_ms[w] = S->comm_stats.total_ms[w] *
                         (double(S->comm_stats.count[w]) / double(S->comm_timed[w]));
    }
  }
  return SLPB_OK;
}

int slpb_get_timers(slpb_solver* S, slpb_timers* out) {
  if (!S || !out) return SLPB_ERR_ARGUMENT;
  CU(cudaSetDevice(S->device));
  CU(cudaStreamSynchronize(S->stream));
  harvest_timers(S);
  *out = S->timers;
  return SLPB_OK;
}

int slpb_flush_l2(slpb_solver* S) {
  if (!S) return SLPB_ERR_ARGUMENT;
  const AllocScope alloc_scope{S->stream};
  CU(cudaSetDevice(S->device));
  constexpr size_t kBytes = size_t(256) << 20;
  if (S->l2_flush.n != kBytes) CU(S->l2_flush.alloc(kBytes));
  CU(cudaMemsetAsync(S->l2_flush.p, 0xA5, kBytes, S->stream));
  CU(cudaStreamSynchronize(S->stream));
  return SLPB_OK;
}

void* slpb_stream(slpb_solver* S) { return S ? S->stream : nullptr; }

}  // extern "C"

#include "batch.cuh"
#include "group.cuh"
