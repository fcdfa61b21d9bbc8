// Warp-specialised front kernels for fronts of order ≤ 32 (device only).
//
// Same assembly and the same per-entry arithmetic as the generic bodies in
// ldlt_core.hpp — W(i,j) ← fma(−l_ik, w_jk, W(i,j)) for k ascending,
// l_ik = w_ik / d_k — so both produce identical bits. What changes is the
// mapping: a front is assembled in shared memory (leading dimension kFrontLd)
// and eliminated in blocks of four pivots, the panel of a block with the rows
// in registers and warp shuffles, the trailing update of dense fronts on the
// FP64 tensor cores (ldlt_dense.cuh); the right-hand side rides along (forward
// substitution fused into the factorisation); the children's update matrices
// are fetched with every load in flight at once and land through a precomputed
// position map; all per-front metadata comes from one packed record instead of
// a chain of dependent index loads.
#pragma once

#include "ldlt_core.hpp"
#include "ldlt_dense.cuh"

namespace slpb {

/// Everything a warp needs to know about a front, in one 80-byte record.
struct alignas(16) FrontMeta {
  int32_t F, np, c0, n_child;
  int32_t child_begin, asm_begin, asm_end, rel_off;
  int64_t panel_off, update_off;
  int64_t rows_off;
  int32_t parent, pad;
  // extend-add list of the warp-per-front factorisation: chunks of 32
  // (source, destination) pairs, see ExtendList
  int32_t ext_begin, ext_chunks, pad2, pad3;
};
static_assert(sizeof(FrontMeta) == 80, "FrontMeta is five 16-byte words");

__device__ __forceinline__ FrontMeta load_front_meta(const FrontMeta* p) {
  // five independent 16-byte loads
  FrontMeta m;
  const int4* src = reinterpret_cast<const int4*>(p);
  int4* dst = reinterpret_cast<int4*>(&m);
  dst[0] = __ldg(src + 0);
  dst[1] = __ldg(src + 1);
  dst[2] = __ldg(src + 2);
  dst[3] = __ldg(src + 3);
  dst[4] = __ldg(src + 4);
  return m;
}

/// Extend-add of a front as one flat list (built once per symbolic analysis):
/// for every child, in child order, the lower triangle of its update matrix
/// and its update vector as (source, destination) pairs, padded per child to a
/// multiple of 32, so that a chunk of 32 pairs belongs to ONE child (its
/// entries land on distinct destinations) and chunks follow the fixed child
/// order of the summation. src: offset into the update-matrix storage, or
/// kExtVecTag | offset into the update-vector storage; dst: offset from the
/// start of the warp's front workspace (the front itself, or the right-hand
/// side behind it); −1 pads.
constexpr int32_t kExtVecTag = 1 << 30;
struct ExtendList {
  const int32_t* src;
  const int32_t* dst;
};
/// Chunks whose loads are all in flight at once.
constexpr int kExtBatch = 8;

/// What a parent front pre-loads about one child before it starts waiting.
struct ChildPre {
  int mc;            // order of the child's update matrix
  int ri;            // target row/col of this lane's row (lane < mc)
  int rel_off;       // child's slice of the update-vector storage
  const double* U;   // child's update matrix
};

__device__ __forceinline__ ChildPre preload_child(
    int lane, int c, const FrontMeta* __restrict__ metas,
    const int32_t* __restrict__ rel_idx, const double* updates) {
  const FrontMeta cm = load_front_meta(metas + c);
  ChildPre p;
  p.mc = cm.F - cm.np;
  p.ri = lane < p.mc ? rel_idx[cm.rel_off + lane] : 0;
  p.rel_off = cm.rel_off;
  p.U = updates + cm.update_off;
  return p;
}

constexpr int kPreChildren = 4;  // children whose metadata is pre-loaded

/// Spins with acquire loads until *p ≥ need (short back-off: the dependency is
/// usually a few hundred nanoseconds away and sits on the critical path).
__device__ __forceinline__ void wait_children(const int* p, int need) {
  if (need <= 0) return;
  int v;
  unsigned spins = 0;
  for (;;) {
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    if (v >= need) break;
    if (++spins > 16) __nanosleep(spins > 256 ? 200 : 40);
  }
}

/// The same wait for a whole warp, WITHOUT a lane-0 branch: lane 0 loads, every
/// lane takes the decision from it, so the warp stays converged (a warp that
/// comes out of a one-lane spin loop split in two runs everything behind it
/// lane group by lane group; measured on the factorisation: 3–4× slower
/// eliminations).
__device__ __forceinline__ void wait_children_warp(int lane, const int* p,
                                                   int need) {
  if (need <= 0) return;
  unsigned spins = 0;
  for (;;) {
    int v = 0;
    if (lane == 0) {
      asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    }
    v = __shfl_sync(0xffffffffu, v, 0);
    if (v >= need) break;
    if (++spins > 16) __nanosleep(spins > 256 ? 200 : 40);
  }
}

/// Right-hand side carried through the factorisation (forward substitution
/// fused into the factor launch); rhs == nullptr switches it off.
struct FusedRhs {
  const double* rhs;      // un-permuted right-hand side
  const int32_t* perm;
  double* xperm;          // y of the own columns, elimination order
  double* uvecs;          // update vectors, indexed like rel_idx
};

/// W: the warp's front workspace (kFrontSmemDoubles of shared memory: the
/// front with leading dimension kFrontLd, the side buffers of the blocked
/// elimination, the front's right-hand side). asm_dst: position of each own
/// KKT entry in that layout; ext: the front's extend-add list. `dep` is polled
/// by lane 0 AFTER everything that does not depend on the children is done —
/// the indices of the first kExtBatch chunks included, so that behind the wait
/// there is ONE round trip to the L2 for the children's data.
__device__ __forceinline__ void ldlt_factor_front_warp(
    int lane, const FrontMeta& fm, const int32_t* __restrict__ asm_src,
    const int32_t* __restrict__ asm_dst, const ExtendList& ext,
    const uint8_t* __restrict__ col_is_primal,
    const double* __restrict__ Kval, double delta, double gamma,
    double* __restrict__ panels, double* updates, double* __restrict__ D,
    double* __restrict__ W, const int* dep, int* local_stats,
    const FusedRhs& fr, bool fused_arith,
    unsigned long long* stamp = nullptr) {
  const int F = fm.F, np = fm.np, c0 = fm.c0, m = F - np;
  double* side = W + kFrontLd * kFrontCols;
  double* col = side + kDenseSideDoubles;  // the front's right-hand side
  // own part of the right-hand side
  if (fr.rhs != nullptr && lane < F) {
    col[lane] = lane < np ? fr.rhs[fr.perm[c0 + lane]] : 0.0;
  }

  // ---- child-independent part: own KKT entries, δ/γ, extend-add indices -----
  for (int j = 0; j < F; ++j) W[lane + j * kFrontLd] = 0.0;
  __syncwarp();
  for (int k = fm.asm_begin + lane; k < fm.asm_end; k += 32) {
    W[asm_dst[k]] = Kval[asm_src[k]];
  }
  const int32_t* esrc = ext.src + size_t(fm.ext_begin) * 32 + lane;
  const int32_t* edst = ext.dst + size_t(fm.ext_begin) * 32 + lane;
  int src[kExtBatch], dst[kExtBatch];
#pragma unroll
  for (int q = 0; q < kExtBatch; ++q) {
    const bool in = q < fm.ext_chunks;
    src[q] = in ? __ldg(esrc + 32 * q) : 0;
    dst[q] = in ? __ldg(edst + 32 * q) : -1;
    // without a right-hand side the update vectors stay where they are
    if (fr.rhs == nullptr && (src[q] & kExtVecTag)) dst[q] = -1;
  }
  __syncwarp();
  if (lane < np) {
    W[lane + lane * kFrontLd] += col_is_primal[c0 + lane] ? delta : -gamma;
  }

  // ---- wait for the children, then extend-add --------------------------------
  wait_children_warp(lane, dep, fm.n_child);
  if (stamp) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    if (lane == 0) *stamp = t;
  }
  __syncwarp();
  {
    double u[kExtBatch];
#pragma unroll
    for (int q = 0; q < kExtBatch; ++q) {
      const double* from = (src[q] & kExtVecTag)
                               ? fr.uvecs + (src[q] & (kExtVecTag - 1))
                               : updates + src[q];
      u[q] = dst[q] >= 0 ? __ldcg(from) : 0.0;
    }
#pragma unroll
    for (int q = 0; q < kExtBatch; ++q) {
      if (q < fm.ext_chunks) {
        if (dst[q] >= 0) W[dst[q]] += u[q];
        __syncwarp();  // the next chunk may belong to the next child
      }
    }
  }
  for (int q0 = kExtBatch; q0 < fm.ext_chunks; q0 += kExtBatch) {
    int sx[kExtBatch], dx[kExtBatch];
    double u[kExtBatch];
#pragma unroll
    for (int q = 0; q < kExtBatch; ++q) {
      const bool in = q0 + q < fm.ext_chunks;
      sx[q] = in ? __ldg(esrc + 32 * (q0 + q)) : 0;
      dx[q] = in ? __ldg(edst + 32 * (q0 + q)) : -1;
      if (fr.rhs == nullptr && (sx[q] & kExtVecTag)) dx[q] = -1;
    }
#pragma unroll
    for (int q = 0; q < kExtBatch; ++q) {
      const double* from = (sx[q] & kExtVecTag)
                               ? fr.uvecs + (sx[q] & (kExtVecTag - 1))
                               : updates + sx[q];
      u[q] = dx[q] >= 0 ? __ldcg(from) : 0.0;
    }
#pragma unroll
    for (int q = 0; q < kExtBatch; ++q) {
      if (q0 + q < fm.ext_chunks) {
        if (dx[q] >= 0) W[dx[q]] += u[q];
        __syncwarp();
      }
    }
  }
  if (stamp) {  // [1] the children were seen, [2] they are added
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    if (lane == 0) stamp[1] = t;
  }
  // (the lanes must ENTER the elimination together: a warp that arrives split,
  // e.g. behind the lane-0 branch above, runs all of it lane group by lane
  // group and only meets at the shuffles — measured 4× slower)
  __syncwarp();
  // ---- elimination of the own columns + write-out ----------------------------
  double rhs_i = (fr.rhs != nullptr && lane < F) ? col[lane] : 0.0;
#ifdef SLPB_DENSE_PROFILE
  const long long call_t0 = clock64();
#endif
  if (fused_arith) {
    ldlt_eliminate_front<true>(lane, F, np, m, W, side, D + c0,
                               panels + fm.panel_off, updates + fm.update_off,
                               local_stats, rhs_i);
  } else {
    ldlt_eliminate_front<false>(lane, F, np, m, W, side, D + c0,
                                panels + fm.panel_off, updates + fm.update_off,
                                local_stats, rhs_i);
  }
#ifdef SLPB_DENSE_PROFILE
  if (lane == 0 && clock64() - call_t0 > 16000) {
    printf("   call F=%d np=%d children=%d: %lld cycles\n", F, np, fm.n_child, clock64() - call_t0);
  }
#endif
  if (fr.rhs != nullptr) {
    if (lane < np) {
      fr.xperm[c0 + lane] = rhs_i;
    } else if (lane < F) {
      fr.uvecs[fm.rel_off + (lane - np)] = rhs_i;
    }
  }
}

/// Forward substitution on a front (same arithmetic as ldlt_forward_front).
/// Lane i keeps w_i and its row of L in registers; everything that does not
/// depend on the children is loaded before the wait.
__device__ __forceinline__ void ldlt_forward_front_warp(
    int lane, const FrontMeta& fm, const FrontMeta* __restrict__ metas,
    const int32_t* __restrict__ child_idx, const int32_t* __restrict__ rel_idx,
    const int32_t* __restrict__ perm, const double* __restrict__ panels,
    const double* __restrict__ rhs, double* x_perm, double* uvecs,
    double* __restrict__ w, const int* dep) {
  const int F = fm.F, np = fm.np, c0 = fm.c0;
  const double* P = panels + fm.panel_off;
  double Lrow[32];  // L(lane, k), k < lane
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    Lrow[k] = (k < np && lane > k && lane < F) ? P[tri_col(k, F) + lane] : 0.0;
  }
  if (lane < F) w[lane] = lane < np ? rhs[perm[c0 + lane]] : 0.0;
  ChildPre pre[kPreChildren];
#pragma unroll
  for (int q = 0; q < kPreChildren; ++q) {
    if (q < fm.n_child) {
      pre[q] = preload_child(lane, child_idx[fm.child_begin + q], metas,
                             rel_idx, nullptr);
    }
  }
  wait_children_warp(lane, dep, fm.n_child);
  __syncwarp();
  for (int ck = 0; ck < fm.n_child; ++ck) {
    ChildPre cp;
    if (ck < kPreChildren) {
      cp = pre[0];
#pragma unroll
      for (int q = 1; q < kPreChildren; ++q) {
        if (ck == q) cp = pre[q];
      }
    } else {
      cp = preload_child(lane, child_idx[fm.child_begin + ck], metas, rel_idx,
                         nullptr);
    }
    if (lane < cp.mc) w[cp.ri] += __ldcg(uvecs + cp.rel_off + lane);
    __syncwarp();
  }
  double wi = lane < F ? w[lane] : 0.0;
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    if (k >= np) break;
    const double yk = __shfl_sync(0xffffffffu, wi, k);
    if (lane > k) wi -= Lrow[k] * yk;
  }
  if (lane < np) {
    x_perm[c0 + lane] = wi;
  } else if (lane < F) {
    uvecs[fm.rel_off + (lane - np)] = wi;
  }
}

/// Backward substitution on a front (same arithmetic as ldlt_backward_front).
/// Lane k keeps column k of L in registers, loaded before the wait.
__device__ __forceinline__ void ldlt_backward_front_warp(
    int lane, const FrontMeta& fm, const int32_t* __restrict__ rows_idx,
    const double* __restrict__ panels, const double* __restrict__ D,
    double* x_perm, const int* dep) {
  const int F = fm.F, np = fm.np, c0 = fm.c0;
  const double* P = panels + fm.panel_off;
  double Lcol[32];  // L(i, lane), i > lane
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    Lcol[i] = (lane < np && i > lane && i < F) ? P[tri_col(lane, F) + i] : 0.0;
  }
  const double dinv_src = lane < np ? D[c0 + lane] : 1.0;
  const int row = (lane >= np && lane < F) ? rows_idx[fm.rows_off + lane] : 0;
  wait_children_warp(lane, dep, 1);
  __syncwarp();
  double wi = 0.0;
  if (lane < np) {
    wi = __ldcg(x_perm + c0 + lane) / dinv_src;
  } else if (lane < F) {
    wi = __ldcg(x_perm + row);
  }
  // t_k = z_k − Σ_{i ≥ np} L(i,k) x_i  (ascending i), then L11ᵀ x = t
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    if (i >= F) break;
    const double xi = __shfl_sync(0xffffffffu, wi, i);
    if (i >= np && lane < np) wi -= Lcol[i] * xi;
  }
#pragma unroll
  for (int i = 31; i >= 1; --i) {
    if (i >= np) continue;
    const double xi = __shfl_sync(0xffffffffu, wi, i);
    if (lane < i) wi -= Lcol[i] * xi;
  }
  if (lane < np) x_perm[c0 + lane] = wi;
}

}  // namespace slpb
