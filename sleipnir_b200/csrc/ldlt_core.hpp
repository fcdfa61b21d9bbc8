// Per-front bodies of the supernodal multifrontal LDLᵀ (factor, forward
// solve, backward solve). They replace Eigen::SimplicialLDLT::factorize /
// solve as the reference uses them (solver/util/sparse_regularized_ldlt.hpp:
// 74,105 and :159-161): lower triangle, NO pivoting, D diagonal, "failure" only
// when a pivot is exactly zero.
//
// Two arithmetic modes, shared by every factorisation path of this library
// (this generic body, the warp-per-front kernels of ldlt_warp.cuh /
// ldlt_dense.cuh and the batched kernels of batch.cuh), so that all paths
// produce the same bits in either mode; l_ik = w_ik / d_k (correctly rounded
// division) in both:
//   reference (default): W(i,j) ← W(i,j) − RN(l_ik · w_jk), product and
//     difference rounded separately like the reference's x86-64 build;
//   fused: every Schur-complement term is ONE fused multiply-add,
//     W(i,j) ← fma(−l_ik, w_jk, W(i,j)), k ascending — what the FP64 tensor
//     core computes (a chain of FMAs over the four pivots of a block,
//     measured), the mode in which dense fronts run on the tensor cores.
// The two differ by one rounding per term, but not harmlessly for the solver
// above: a structurally singular pivot, mathematically zero, comes out EXACTLY
// zero under separate rounding in most cases (RN(RN(w/d)·w) lands back on the
// entry it is subtracted from) and is then caught by the zero-pivot / inertia
// test of the regularisation loop, while the fused update leaves a residual of
// ~1e-17 that can pass for a legitimate pivot. Whole solves therefore follow
// the reference's decisions only in the reference mode.
//
// One front = one supernode of the assembly tree (symbolic.cpp): a dense
// F×F column-major matrix whose first `np` rows/columns are the supernode's own
// (permuted) columns. Written once as functions of (tid, NT): the kernels run
// them with one thread block per front (front in shared memory, __syncthreads
// between phases); tests/emu runs the same code on the host with NT = 1.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>

#include "ad_core.hpp"  // SLPB_HD

namespace slpb {

/// Packed lower triangle of order n, column-major: offset of column j minus j,
/// so that entry (i, j), i ≥ j, sits at tri_col(j, n) + i. The L panel of a
/// front keeps the first np columns of the triangle of order F (its diagonal
/// slots hold d), the update matrix the triangle of order m = F − np.
SLPB_HD int tri_col(int j, int n) {
  return j * n - (j * (j - 1)) / 2 - j;
}

/// Device-side view of struct Symbolic (plain pointers).
struct SymbolicView {
  int32_t dim, n_super;
  const int32_t* super_first;
  const int32_t* front_dim;
  const int64_t* rows_ptr;
  const int32_t* rows_idx;
  const int64_t* panel_ptr;
  const int64_t* update_ptr;
  const int64_t* child_ptr;
  const int32_t* child_idx;
  const int64_t* rel_ptr;
  const int32_t* rel_idx;
  const int64_t* asm_ptr;
  const int32_t* asm_src;
  const int32_t* asm_dst;
  const uint8_t* col_is_primal;
  const int32_t* perm;
};

/// Inertia bookkeeping of one factorisation (device memory).
struct FactorStats {
  int32_t n_pos, n_neg, n_zero, zero_pivot;
  unsigned long long min_abs_d_bits;  // bit pattern of min |D_ii| (positive)
};

/// Assembles and partially factors front `s`. W: F*F doubles of scratch,
/// lcol: F doubles of scratch. Writes the L panel, D and the update matrix.
template <int NT, typename Sync>
SLPB_HD void ldlt_factor_front(int tid, int s, const SymbolicView& S,
                               const double* __restrict__ Kval, double delta,
                               double gamma, double* __restrict__ panels,
                               double* __restrict__ updates,
                               double* __restrict__ D, double* W, double* lcol,
                               int* local_stats, Sync sync, bool fused = false) {
  const int F = S.front_dim[s];
  const int c0 = S.super_first[s];
  const int np = S.super_first[s + 1] - c0;
  const int m = F - np;

  for (int i = tid; i < F * F; i += NT) W[i] = 0.0;
  sync();
  // scatter the KKT entries owned by this supernode
  for (int64_t k = S.asm_ptr[s] + tid; k < S.asm_ptr[s + 1]; k += NT) {
    W[S.asm_dst[k]] = Kval[S.asm_src[k]];
  }
  sync();
  // + diag(δ…δ, −γ…−γ) on the own columns
  for (int j = tid; j < np; j += NT) {
    W[j + j * F] += S.col_is_primal[c0 + j] ? delta : -gamma;
  }
  sync();
  // extend-add the children's update matrices, one child after the other so
  // that every entry is accumulated in a fixed order
  for (int64_t ck = S.child_ptr[s]; ck < S.child_ptr[s + 1]; ++ck) {
    const int c = S.child_idx[ck];
    const int mc = S.front_dim[c] - (S.super_first[c + 1] - S.super_first[c]);
    const double* U = updates + S.update_ptr[c];
    const int32_t* rel = S.rel_idx + S.rel_ptr[c];
    {
      constexpr int LW = NT >= 32 ? 32 : NT;
      constexpr int NG = NT / LW;
      for (int j = tid / LW; j < mc; j += NG) {
        const int cj = rel[j] * F;
        const double* Uj = U + tri_col(j, mc);
        for (int i = j + tid % LW; i < mc; i += LW) {
          W[rel[i] + cj] += SLPB_LDCG(Uj + i);
        }
      }
    }
    sync();
  }
  // right-looking elimination of the own columns
  int pos = 0, neg = 0, zero = 0, zpiv = 0;
  double min_abs = INFINITY;
  for (int k = 0; k < np; ++k) {
    const double d = W[k + k * F];
    if (tid == 0) {
      const double eps = 2.220446049250313e-16;
      if (d > eps) {
        ++pos;
      } else if (d < -eps) {
        ++neg;
      } else {
        ++zero;
      }
      if (d == 0.0) zpiv = 1;
      min_abs = fmin(min_abs, fabs(d));
      D[c0 + k] = d;
    }
    for (int i = k + 1 + tid; i < F; i += NT) lcol[i] = W[i + k * F] / d;
    sync();
    // W(i,j) −= l_ik · (d·l_jk) for k < j ≤ i, with d·l_jk still in column k.
    // Threads are laid out two-dimensionally — groups of up to 32 take a column
    // each, the threads of a group its rows — so that a big front keeps a whole
    // thread block busy (every entry still receives exactly this one update
    // per pivot: the mapping does not change any bit).
    {
      constexpr int LW = NT >= 32 ? 32 : NT;
      constexpr int NG = NT / LW;
      const int lane_in_group = tid % LW, group = tid / LW;
      for (int j = k + 1 + group; j < F; j += NG) {
        const double wjk = W[j + k * F];
        if (fused) {
          for (int i = j + lane_in_group; i < F; i += LW) {
            W[i + j * F] = fma(-lcol[i], wjk, W[i + j * F]);
          }
        } else {
          for (int i = j + lane_in_group; i < F; i += LW) {
            W[i + j * F] -= lcol[i] * wjk;
          }
        }
      }
    }
    sync();
    for (int i = k + 1 + tid; i < F; i += NT) W[i + k * F] = lcol[i];
    sync();
  }
  // L panel (F × np, unit diagonal implicit) and update matrix (m × m)
  double* P = panels + S.panel_ptr[s];
  for (int k = 0; k < np; ++k) {
    double* Pk = P + tri_col(k, F);
    for (int i = k + tid; i < F; i += NT) Pk[i] = W[i + k * F];
  }
  double* U = updates + S.update_ptr[s];
  for (int j = 0; j < m; ++j) {
    double* Uj = U + tri_col(j, m);
    for (int i = j + tid; i < m; i += NT) Uj[i] = W[(np + i) + (np + j) * F];
  }
  if (tid == 0) {
    local_stats[0] = pos;
    local_stats[1] = neg;
    local_stats[2] = zero;
    local_stats[3] = zpiv;
    // |d| ≥ 0: IEEE bit patterns of non-negative doubles order like integers
    double a = min_abs;
    unsigned long long bits;
    memcpy(&bits, &a, 8);
    local_stats[4] = static_cast<int>(bits & 0xffffffffull);
    local_stats[5] = static_cast<int>(bits >> 32);
  }
}

/// Forward substitution on front s: w = [b_own + Σ children; Σ children],
/// solve L11 y = w_own, u = w_below − L21 y. x_perm[own] ← y, uvec[s] ← u.
template <int NT, typename Sync>
SLPB_HD void ldlt_forward_front(int tid, int s, const SymbolicView& S,
                                const double* __restrict__ panels,
                                const double* __restrict__ rhs,
                                double* __restrict__ x_perm,
                                double* __restrict__ uvecs, double* w,
                                Sync sync) {
  const int F = S.front_dim[s];
  const int c0 = S.super_first[s];
  const int np = S.super_first[s + 1] - c0;
  for (int i = tid; i < F; i += NT) {
    w[i] = i < np ? rhs[S.perm[c0 + i]] : 0.0;
  }
  sync();
  for (int64_t ck = S.child_ptr[s]; ck < S.child_ptr[s + 1]; ++ck) {
    const int c = S.child_idx[ck];
    const int mc = S.front_dim[c] - (S.super_first[c + 1] - S.super_first[c]);
    const double* u = uvecs + S.rel_ptr[c];
    const int32_t* rel = S.rel_idx + S.rel_ptr[c];
    for (int i = tid; i < mc; i += NT) w[rel[i]] += SLPB_LDCG(u + i);
    sync();
  }
  const double* P = panels + S.panel_ptr[s];
  for (int k = 0; k < np; ++k) {
    const double yk = w[k];
    const double* Pk = P + tri_col(k, F);
    for (int i = k + 1 + tid; i < F; i += NT) w[i] -= Pk[i] * yk;
    sync();
  }
  for (int i = tid; i < F; i += NT) {
    if (i < np) {
      x_perm[c0 + i] = w[i];
    } else {
      uvecs[S.rel_ptr[s] + (i - np)] = w[i];
    }
  }
}

/// Backward substitution on front s (after its ancestors): z = y_own / d,
/// x_own = L11⁻ᵀ (z − L21ᵀ x_below).
template <int NT, typename Sync>
SLPB_HD void ldlt_backward_front(int tid, int s, const SymbolicView& S,
                                 const double* __restrict__ panels,
                                 const double* __restrict__ D,
                                 double* __restrict__ x_perm, double* w,
                                 Sync sync) {
  const int F = S.front_dim[s];
  const int c0 = S.super_first[s];
  const int np = S.super_first[s + 1] - c0;
  const int32_t* rows = S.rows_idx + S.rows_ptr[s];
  for (int i = tid; i < F; i += NT) {
    w[i] = i < np ? SLPB_LDCG(x_perm + c0 + i) / D[c0 + i]
                  : SLPB_LDCG(x_perm + rows[i]);
  }
  sync();
  const double* P = panels + S.panel_ptr[s];
  // t = z − L21ᵀ x_below: one own column per thread
  for (int k = tid; k < np; k += NT) {
    double acc = w[k];
    const double* Pk = P + tri_col(k, F);
    for (int i = np; i < F; ++i) acc -= Pk[i] * w[i];
    w[k] = acc;
  }
  sync();
  // L11ᵀ x = t, column-oriented: once x_i is final, every k < i takes its term
  for (int i = np - 1; i >= 1; --i) {
    const double xi = w[i];
    for (int k = tid; k < i; k += NT) w[k] -= P[tri_col(k, F) + i] * xi;
    sync();
  }
  for (int i = tid; i < np; i += NT) x_perm[c0 + i] = w[i];
}

}  // namespace slpb
