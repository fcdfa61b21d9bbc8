// Warp-specialised front kernels for fronts of order ≤ 32 (device only).
//
// Same assembly and the same per-entry arithmetic as the generic bodies in
// ldlt_core.hpp — W(i,j) −= l_ik · w_jk for k ascending, l_ik = w_ik / d_k — so
// both produce identical bits. What changes is the mapping: a front is assembled
// in shared memory, then lane i takes row i into registers and the pivots are
// eliminated with warp shuffles (no shared-memory round trip, no barrier per
// pivot); the right-hand side rides along (forward substitution fused into the
// factorisation); global loads are batched ahead of the shared-memory
// read-modify-writes, and all per-front metadata comes from one packed record
// instead of a chain of dependent index loads.
#pragma once

#include "ldlt_core.hpp"

namespace slpb {

/// Everything a warp needs to know about a front, in one 64-byte record.
struct alignas(16) FrontMeta {
  int32_t F, np, c0, n_child;
  int32_t child_begin, asm_begin, asm_end, rel_off;
  int64_t panel_off, update_off;
  int64_t rows_off;
  int32_t parent, pad;
};
static_assert(sizeof(FrontMeta) == 64, "FrontMeta must stay one 64-byte line");

__device__ __forceinline__ FrontMeta load_front_meta(const FrontMeta* p) {
  // four independent 16-byte loads
  FrontMeta m;
  const int4* src = reinterpret_cast<const int4*>(p);
  int4* dst = reinterpret_cast<int4*>(&m);
  dst[0] = __ldg(src + 0);
  dst[1] = __ldg(src + 1);
  dst[2] = __ldg(src + 2);
  dst[3] = __ldg(src + 3);
  return m;
}

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

/// Elimination of the own columns of a front of order F ≤ 32 with the rows in
/// registers, then the write-out of the L panel and the update matrix.
/// Lane i keeps row i, shifted so that w[0] is always the current pivot column
/// (w[jj] = W(i, k + jj)): one compact loop body serves every pivot (a fully
/// unrolled 32 × 32 elimination does not fit the instruction cache). Pivot k:
/// lane k broadcasts d, every lane i > k forms l_ik = w_ik / d, the unscaled
/// w_jk travels from lane j by shuffle and W(i,j) −= l_ik · w_jk — the same
/// products in the same order as ldlt_factor_front, with no shared-memory
/// round trip and no warp barrier per pivot. Inertia counters are meaningful
/// in lane 0 only.
__device__ __noinline__ void ldlt_eliminate_rows(
    int lane, int F, int np, int m, const double* __restrict__ W,
    double* __restrict__ Dk, double* __restrict__ P, double* __restrict__ U,
    int* __restrict__ local_stats, double& rhs_i) {
  // rhs_i: this lane's entry of the front's right-hand side; it takes part in
  // the elimination (forward substitution carried along: r_i −= l_ik · r_k,
  // the arithmetic of ldlt_forward_front) and returns y (lanes < np) and the
  // update vector (lanes ≥ np).
  double r = rhs_i;
  double w[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    w[j] = (lane < F && j <= lane) ? W[lane + j * F] : 0.0;
  }
  int pos = 0, neg = 0, zero = 0, zpiv = 0;
  double min_abs = INFINITY;
#pragma unroll 1
  for (int k = 0; k < np; ++k) {
    const double d = __shfl_sync(0xffffffffu, w[0], k);
    {
      // inertia bookkeeping (every lane computes it; lane 0 reports it)
      const double eps = 2.220446049250313e-16;
      pos += d > eps ? 1 : 0;
      neg += d < -eps ? 1 : 0;
      zero += (d > eps || d < -eps) ? 0 : 1;
      zpiv |= d == 0.0 ? 1 : 0;
      min_abs = fmin(min_abs, fabs(d));
    }
    if (lane == 0) Dk[k] = d;
    const double wk = w[0];
    // 0 / d: div.rn.f64 sends a zero dividend through its ~90-instruction
    // special-case subroutine, and every lane outside the column (and every
    // structural zero inside it) would drag the warp through it on every
    // pivot. The quotient is a signed zero: form it directly (same bits).
    double l;
    if (wk == 0.0 && d == d && d != 0.0) {
      l = (signbit(wk) != signbit(d)) ? -0.0 : 0.0;
    } else {
      l = wk / d;
    }
    // column k of the L panel (the diagonal slot keeps d, as in the generic body)
    if (lane < F) P[lane + k * F] = lane > k ? l : wk;
    {
      const double rk = __shfl_sync(0xffffffffu, r, k);
      const double upd = r - l * rk;
      r = (lane > k && lane < F) ? upd : r;
    }
    // trailing update in branch-free groups of 8 columns, so that the eight
    // shuffles and multiply-subtracts of a group overlap
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      if (k + 1 + 8 * g < F) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int jj = 1 + 8 * g + q;
          if (jj < 32) {
            const int j = k + jj;
            const double wjk = __shfl_sync(0xffffffffu, wk, j & 31);
            const double upd = w[jj] - l * wjk;
            w[jj] = (j < F && lane >= j) ? upd : w[jj];
          }
        }
      }
    }
#pragma unroll
    for (int jj = 0; jj < 31; ++jj) w[jj] = w[jj + 1];
    w[31] = 0.0;
  }
  // update matrix (m × m, lower): w[jj] now holds column np + jj
#pragma unroll
  for (int jj = 0; jj < 32; ++jj) {
    if (jj < m && lane >= np + jj && lane < F) {
      U[(lane - np) + jj * m] = w[jj];
    }
  }
  rhs_i = r;
  if (lane == 0) {
    local_stats[0] = pos;
    local_stats[1] = neg;
    local_stats[2] = zero;
    local_stats[3] = zpiv;
    const unsigned long long bits = __double_as_longlong(min_abs);
    local_stats[4] = static_cast<int>(bits & 0xffffffffull);
    local_stats[5] = static_cast<int>(bits >> 32);
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

/// W: F×F column-major in shared memory; col: 64 doubles of shared scratch (the
/// front's right-hand side while it is assembled). `dep` is polled by lane 0
/// AFTER everything that does not depend on the children is done.
__device__ __forceinline__ void ldlt_factor_front_warp(
    int lane, const FrontMeta& fm, const FrontMeta* __restrict__ metas,
    const int32_t* __restrict__ child_idx, const int32_t* __restrict__ rel_idx,
    const int32_t* __restrict__ asm_src, const int32_t* __restrict__ asm_dst,
    const uint8_t* __restrict__ col_is_primal,
    const double* __restrict__ Kval, double delta, double gamma,
    double* __restrict__ panels, double* updates, double* __restrict__ D,
    double* __restrict__ W, double* __restrict__ col,
    const int* dep, int* local_stats, const FusedRhs& fr,
    unsigned long long* stamp = nullptr) {
  const int F = fm.F, np = fm.np, c0 = fm.c0, m = F - np;
  // own part of the right-hand side (col[] doubles as the rhs work vector)
  if (fr.rhs != nullptr && lane < F) {
    col[lane] = lane < np ? fr.rhs[fr.perm[c0 + lane]] : 0.0;
  }

  // ---- child-independent part: own KKT entries, δ/γ, child metadata ---------
  for (int i = lane; i < F * F; i += 32) W[i] = 0.0;
  __syncwarp();
  for (int k = fm.asm_begin + lane; k < fm.asm_end; k += 32) {
    W[asm_dst[k]] = Kval[asm_src[k]];
  }
  ChildPre pre[kPreChildren];
#pragma unroll
  for (int q = 0; q < kPreChildren; ++q) {
    if (q < fm.n_child) {
      pre[q] = preload_child(lane, child_idx[fm.child_begin + q], metas,
                             rel_idx, updates);
    }
  }
  __syncwarp();
  if (lane < np) W[lane + lane * F] += col_is_primal[c0 + lane] ? delta : -gamma;

  // ---- wait for the children, then extend-add their update matrices ---------
  if (lane == 0) {
    wait_children(dep, fm.n_child);
    if (stamp) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
      *stamp = t;
    }
  }
  __syncwarp();
  for (int ck = 0; ck < fm.n_child; ++ck) {
    ChildPre cp;
    if (ck < kPreChildren) {
      // static indexing keeps pre[] in registers
      cp = pre[0];
#pragma unroll
      for (int q = 1; q < kPreChildren; ++q) {
        if (ck == q) cp = pre[q];
      }
    } else {
      cp = preload_child(lane, child_idx[fm.child_begin + ck], metas, rel_idx,
                         updates);
    }
    const int mc = cp.mc, ri = cp.ri;
    const double* U = cp.U;
    if (fr.rhs != nullptr && lane < mc) {
      col[ri] += __ldcg(fr.uvecs + cp.rel_off + lane);
    }
    for (int j0 = 0; j0 < mc; j0 += 8) {
      double u[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int j = j0 + q;
        u[q] = (j < mc && j <= lane && lane < mc) ? __ldcg(U + lane + j * mc)
                                                  : 0.0;
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int j = j0 + q;
        const int rj = __shfl_sync(0xffffffffu, ri, j & 31);
        if (j < mc && j <= lane && lane < mc) W[ri + rj * F] += u[q];
      }
    }
    __syncwarp();
  }

  // ---- elimination of the own columns + write-out (separate function so that
  // the 32-double row stays in registers) --------------------------------------
  double rhs_i = (fr.rhs != nullptr && lane < F) ? col[lane] : 0.0;
  ldlt_eliminate_rows(lane, F, np, m, W, D + c0, panels + fm.panel_off,
                      updates + fm.update_off, local_stats, rhs_i);
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
    Lrow[k] = (k < np && lane > k && lane < F) ? P[lane + k * F] : 0.0;
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
  if (lane == 0) wait_children(dep, fm.n_child);
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
    Lcol[i] = (lane < np && i > lane && i < F) ? P[i + lane * F] : 0.0;
  }
  const double dinv_src = lane < np ? D[c0 + lane] : 1.0;
  const int row = (lane >= np && lane < F) ? rows_idx[fm.rows_off + lane] : 0;
  if (lane == 0) wait_children(dep, 1);
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
