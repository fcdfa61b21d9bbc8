// Many KKT systems over ONE symbolic structure, factored and solved side by
// side with lane = instance (included at the end of slpb.cu).
//
// Reference surface: slp::multistart (optimization/multistart.hpp:44-73) solves
// the same problem from many initial guesses; every start has the lhs pattern
// of interior_point.hpp:426-440 and differs only in values. One numeric LDLᵀ of
// that system is a 13-level dependency chain over a factor that fits L2 many
// times over: latency-bound, ≈1 % of the HBM roofline. B
// systems in SoA layout walk the SAME chain once and move B× the bytes — the
// regime in which the factorisation and the triangular solves are bound by
// memory traffic.
//
// Layout. Instances come in groups of 32; every per-entry array of a group is
// stored [entry][32], so the 32 lanes of a warp read one 256-byte line per
// entry (fully coalesced), and group g of an array with E entries starts at
// g·E·32. A warp owns one (front, group) task: lane ℓ runs the scalar
// elimination of ldlt_core.hpp (ldlt_factor_front with NT = 1) for instance
// 32·g + ℓ — the same products in the same order, hence bit-identical D, L and
// solutions to the single-instance kernels — on a frontal matrix whose lower
// triangle sits interleaved in shared memory (W[e·32 + ℓ], conflict-free, no
// cross-lane traffic and no barrier inside a front). Tasks are drawn from a
// ticket counter in level order, children before parents, and hand over through
// per-(group, front) counters exactly like k_factor_tree.
//
// Storage is packed: a front keeps the lower trapezoid of its panel
// (np·F − np(np−1)/2 entries, column-major) and the lower triangle of its update
// matrix (m(m+1)/2) — in the column-major packed lower triangle of the F×F front
// the two are contiguous, so the write-out is one flat copy.
#pragma once

namespace slpb {

constexpr int kBatchLanes = 32;
// Pivots applied to the trailing matrix at once, and warps that share one
// (front, group) task: all of them work for the SAME 32 instances (lane =
// instance) and split the rows / columns of the front between them.
// (measured on B200, cart-pole N=5000, B=512: rank 2 / 6 warps 3.47 ms,
// rank 2 / 8 warps 3.67, rank 2 / 4 warps 3.76, rank 4 / 8 warps 4.7 — rank 4
// halves the shared-memory traffic per update but its larger column buffer
// leaves room for two blocks per SM instead of three)
#ifndef SLPB_BATCH_RANK
#define SLPB_BATCH_RANK 2
#endif
#ifndef SLPB_BATCH_WARPS
#define SLPB_BATCH_WARPS 6
#endif
#ifndef SLPB_BATCH_MIN_BLOCKS
#define SLPB_BATCH_MIN_BLOCKS (SLPB_BATCH_RANK <= 2 ? 3 : 2)
#endif
constexpr int kBatchRank = SLPB_BATCH_RANK;
constexpr int kBatchWarps = SLPB_BATCH_WARPS;
/// Barrier between the phases of a task.
__device__ __forceinline__ void compute_sync() { __syncthreads(); }

struct BatchView {
  const int32_t* order;      // fronts by ascending level
  const FrontMeta* metas;    // panel_off / update_off: PACKED offsets (entries)
  const int32_t* child_idx;
  const int32_t* rel_idx;
  const int32_t* rows_idx;
  const int32_t* asm_src;    // KKT entry of each own value
  const int32_t* asm_tri;    // its packed position in the front | diagonal flags
  const int32_t* umap;       // packed update entry → packed position in parent
  const uint8_t* col_is_primal;
  const int32_t* perm;
  int32_t* sync;             // [0] ticket | fcount[G·ns] | fflag[G·ns] | bflag[G·ns]
  int32_t n_super, groups, dim, tri_cap;
  int64_t nK, panel_total, update_total, rel_total;
};

/// asm_tri entries: packed position in the low bits; the own diagonal entries
/// carry which regularisation they take.
constexpr int32_t kAsmPrimalDiag = 1 << 30, kAsmDualDiag = 1 << 29;
constexpr int32_t kAsmIndexMask = (1 << 29) - 1;

/// Reciprocal of a pivot, correctly rounded (computed once per pivot and reused
/// by every division of its column).
__device__ __forceinline__ double batch_rcp(double d) { return __drcp_rn(d); }

/// w / d, correctly rounded like div.rn.f64 (so that the batch stays
/// bit-identical to the single-instance kernels and to the reference's
/// x86-64 division), from the correctly rounded reciprocal y = RN(1/d):
/// q0 = RN(w·y) is within two ulps of the quotient; r = w − q·d is exact in one
/// FMA, and q ← RN(q + r·y) (Markstein's division step) first makes q faithful
/// and then, applied once more, correctly rounded. A column costs ONE
/// reciprocal and five FP64 operations per entry instead of a full division
/// (≈30 instructions) per entry. Whenever a lane falls outside the range in
/// which the steps are exact (subnormal, huge or non-finite quotient or pivot)
/// the whole warp takes div.rn.f64. tests/test_division_step.py checks the
/// sequence against exact rational arithmetic.
__device__ __forceinline__ double batch_div(double w, double d, double y) {
  // biased exponent in [123, 1923] ⇔ 2^-900 ≤ |x| < 2^901
  const unsigned ed = (static_cast<unsigned>(__double2hiint(d)) >> 20) & 0x7ffu;
  const bool d_ok = ed - 123u <= 1800u;
  // a structural zero of the column (zero in every lane): the signed zero
  // div.rn.f64 returns, no arithmetic at all
  if (__all_sync(0xffffffffu, w == 0.0 && d_ok)) {
    return (signbit(w) != signbit(d)) ? -0.0 : 0.0;
  }
  // (these must stay fused even though the file is built with -fmad=false)
  double q = __dmul_rn(w, y);
  q = __fma_rn(__fma_rn(-q, d, w), y, q);
  q = __fma_rn(__fma_rn(-q, d, w), y, q);
  const unsigned eq = (static_cast<unsigned>(__double2hiint(q)) >> 20) & 0x7ffu;
  // (a lane whose dividend alone is zero, or any value out of range, sends
  // the warp to the full division)
  if (__all_sync(0xffffffffu, d_ok && eq - 123u <= 1800u)) return q;
  return w / d;
}

/// One Schur-complement term in the two arithmetic modes of ldlt_core.hpp:
/// c − RN(l·w) (reference) or fma(−l, w, c) (fused / tensor).
template <bool kFused>
__device__ __forceinline__ double schur_term(double c, double l, double w) {
  return kFused ? __fma_rn(-l, w, c) : __dsub_rn(c, __dmul_rn(l, w));
}

/// One (front, group) task of the batched factorisation, run by the kBatchWarps
/// warps of a thread block: every warp works on the same 32 instances (lane =
/// instance) and takes the entries / rows / columns congruent to its index.
/// W: packed lower triangle [n_tri][32] (shared memory, or global scratch for
/// the rare front above the shared-memory cap), lbuf: [kBatchRank][F][32] scaled
/// columns of the current pivot block.
///
/// Pivots go in blocks of up to kBatchRank. Phase A (one barrier): every warp
/// eliminates the small diagonal block redundantly in registers, then takes
/// its rows below the block through the block's pivots — updated unscaled
/// entries back to W (the trailing update reads them as w_jk), scaled ones to
/// lbuf and straight to the packed panel in global memory. Phase B (one
/// barrier): the trailing columns, split over the warps, take all pivots of
/// the block at once. Per entry the operations and their order are those of
/// ldlt_factor_front in the same arithmetic mode (kFused): W(i,j) −= l_ik·w_jk
/// for k ascending, l_ik = w_ik / d_k.
/// kShared: W and lbuf are shared memory (the compiler must see that to emit
/// LDS/STS with 32-bit addresses instead of generic loads).
template <bool kShared, bool kFused>
__device__ __forceinline__ void batch_factor_front(
    int warp, int lane, const FrontMeta& fm, const BatchView& T, int g,
    const double* __restrict__ Kb, double delta, double gamma,
    double* __restrict__ Pb, double* Ub, double* __restrict__ Db, double* W,
    double* lbuf, const int* dep, int* stats_out /*[6], warp 0 only*/) {
  constexpr int NW = kBatchWarps;
  constexpr int R = kBatchRank;
  const int F = fm.F, np = fm.np, c0 = fm.c0;
  const int n_tri = F * (F + 1) / 2;
  const int n_panel = np * F - (np * (np - 1)) / 2;
  double* Wl = W + lane;
  // ---- own KKT entries (+δ / −γ on the own diagonal) -------------------------
  // W is all zeros here (every task leaves it so): only the entries that
  // have a source are touched, four value loads in flight per warp.
  {
    const double* Kg = Kb + int64_t(g) * T.nK * 32 + lane;
    int k = fm.asm_begin + warp * 4;
    for (; k < fm.asm_end; k += 4 * NW) {
      double v[4];
      int dst[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const bool in = k + q < fm.asm_end;
        dst[q] = in ? __ldg(T.asm_tri + k + q) : 0;
        v[q] = in ? __ldg(Kg + int64_t(__ldg(T.asm_src + k + q)) * 32) : 0.0;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (k + q < fm.asm_end) {
          if (dst[q] & kAsmPrimalDiag) v[q] += delta;
          if (dst[q] & kAsmDualDiag) v[q] += -gamma;
          Wl[(dst[q] & kAsmIndexMask) * 32] = v[q];
        }
      }
    }
  }
  // ---- children: wait, then extend-add in child order ------------------------
  // (warp 0 waits as a whole, without a one-lane branch: see wait_children_warp)
  if (warp == 0) wait_children_warp(lane, dep, fm.n_child);
  compute_sync();
  for (int ck = 0; ck < fm.n_child; ++ck) {
    const FrontMeta cm = load_front_meta(T.metas + T.child_idx[fm.child_begin + ck]);
    const int mc = cm.F - cm.np;
    const int nt = mc * (mc + 1) / 2;
    const double* U =
        Ub + (int64_t(g) * T.update_total + cm.update_off) * 32 + lane;
    const int32_t* map = T.umap + cm.update_off;
    // entries of ONE child land on distinct parent entries: the warps split
    // them freely; the barrier keeps the children in order
    for (int e = warp * 4; e < nt; e += 4 * NW) {
      double u[4];
      int dst[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const bool in = e + q < nt;
        u[q] = in ? __ldcg(U + int64_t(e + q) * 32) : 0.0;
        dst[q] = in ? __ldg(map + e + q) : 0;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (e + q < nt) Wl[dst[q] * 32] += u[q];
      }
    }
    compute_sync();
  }
  // ---- elimination of the own columns ----------------------------------------
  int pos = 0, neg = 0, zero = 0, zpiv = 0;
  double min_abs = INFINITY;
  double* Dg = Db + (int64_t(g) * T.dim + c0) * 32 + lane;
  double* Pg = Pb + (int64_t(g) * T.panel_total + fm.panel_off) * 32 + lane;
  double* Ll = lbuf + lane;
  for (int k0 = 0; k0 < np; k0 += R) {
    const int kb = min(R, np - k0);
    const int jend = k0 + kb;  // first row/column behind the block
    int cb[R];                 // tri_col of the block's columns
#pragma unroll
    for (int q = 0; q < R; ++q) cb[q] = tri_col(min(k0 + q, F - 1), F);
    // -- phase A.1: the kb×kb diagonal block, redundantly in every warp --------
    double a[R][R], ls[R][R], dd[R], ry[R];
#pragma unroll
    for (int q = 0; q < R; ++q) {
#pragma unroll
      for (int p = q; p < R; ++p) {
        a[p][q] = p < kb ? Wl[(cb[q] + k0 + p) * 32] : 0.0;
      }
    }
#pragma unroll
    for (int q = 0; q < R; ++q) {
      dd[q] = 1.0;
      ry[q] = 1.0;
      if (q < kb) {
        dd[q] = a[q][q];
        ry[q] = batch_rcp(dd[q]);
#pragma unroll
        for (int p = q + 1; p < R; ++p) {
          if (p < kb) {
            ls[p][q] = batch_div(a[p][q], dd[q], ry[q]);
#pragma unroll
            for (int q2 = q + 1; q2 <= p; ++q2) {
              a[p][q2] = schur_term<kFused>(a[p][q2], ls[p][q], a[q2][q]);
            }
          }
        }
      }
    }
    if (warp == 0) {
#pragma unroll
      for (int q = 0; q < R; ++q) {
        if (q < kb) {
          const double d = dd[q];
          const double eps = 2.220446049250313e-16;
          pos += d > eps ? 1 : 0;
          neg += d < -eps ? 1 : 0;
          zero += (d > eps || d < -eps) ? 0 : 1;
          zpiv |= d == 0.0 ? 1 : 0;
          min_abs = fmin(min_abs, fabs(d));
          Dg[(k0 + q) * 32] = d;
          Pg[int64_t(cb[q] + k0 + q) * 32] = d;
#pragma unroll
          for (int p = q + 1; p < R; ++p) {
            if (p < kb) Pg[int64_t(cb[q] + k0 + p) * 32] = ls[p][q];
          }
        }
      }
    }
    // -- phase A.2: rows below the block, two at a time per warp ---------------
    for (int i0 = jend + warp; i0 < F; i0 += 2 * NW) {
      const int i1 = i0 + NW;
      const bool two = i1 < F;
      double w0[R], w1[R], l0[R], l1[R];
#pragma unroll
      for (int q = 0; q < R; ++q) {
        w0[q] = q < kb ? Wl[(cb[q] + i0) * 32] : 0.0;
        w1[q] = (q < kb && two) ? Wl[(cb[q] + i1) * 32] : 0.0;
      }
#pragma unroll
      for (int q = 0; q < R; ++q) {
        if (q < kb) {
#pragma unroll
          for (int q1 = 0; q1 < q; ++q1) {
            w0[q] = schur_term<kFused>(w0[q], l0[q1], a[q][q1]);
            w1[q] = schur_term<kFused>(w1[q], l1[q1], a[q][q1]);
          }
          l0[q] = batch_div(w0[q], dd[q], ry[q]);
          l1[q] = batch_div(w1[q], dd[q], ry[q]);
        }
      }
#pragma unroll
      for (int q = 0; q < R; ++q) {
        if (q < kb) {
          Wl[(cb[q] + i0) * 32] = w0[q];
          Ll[(q * F + i0) * 32] = l0[q];
          Pg[int64_t(cb[q] + i0) * 32] = l0[q];
          if (two) {
            Wl[(cb[q] + i1) * 32] = w1[q];
            Ll[(q * F + i1) * 32] = l1[q];
            Pg[int64_t(cb[q] + i1) * 32] = l1[q];
          }
        }
      }
    }
    if (jend >= F) break;  // nothing behind the block
    compute_sync();
    // -- phase B: trailing columns, all pivots of the block at once ------------
    for (int j = jend + warp; j < F; j += NW) {
      double wj[R];
#pragma unroll
      for (int q = 0; q < R; ++q) wj[q] = q < kb ? Wl[(cb[q] + j) * 32] : 0.0;
      double* Wj = Wl + tri_col(j, F) * 32;
      if (kb == R) {
        int i = j;
        for (; i + 2 <= F; i += 2) {
          double x0 = Wj[i * 32], x1 = Wj[(i + 1) * 32];
          double m0[R], m1[R];
#pragma unroll
          for (int q = 0; q < R; ++q) {
            m0[q] = Ll[(q * F + i) * 32];
            m1[q] = Ll[(q * F + i + 1) * 32];
          }
#pragma unroll
          for (int q = 0; q < R; ++q) {
            x0 = schur_term<kFused>(x0, m0[q], wj[q]);
            x1 = schur_term<kFused>(x1, m1[q], wj[q]);
          }
          Wj[i * 32] = x0;
          Wj[(i + 1) * 32] = x1;
        }
        for (; i < F; ++i) {
          double x = Wj[i * 32];
#pragma unroll
          for (int q = 0; q < R; ++q) {
            x = schur_term<kFused>(x, Ll[(q * F + i) * 32], wj[q]);
          }
          Wj[i * 32] = x;
        }
      } else {
        for (int i = j; i < F; ++i) {
          double x = Wj[i * 32];
#pragma unroll
          for (int q = 0; q < R; ++q) {
            if (q < kb) x = schur_term<kFused>(x, Ll[(q * F + i) * 32], wj[q]);
          }
          Wj[i * 32] = x;
        }
      }
    }
    compute_sync();
  }
  // ---- update matrix: the tail of the packed triangle; W goes back to zero ----
  {
    for (int e = warp; e < n_panel; e += NW) Wl[e * 32] = 0.0;
    double* Ug = Ub + (int64_t(g) * T.update_total + fm.update_off) * 32 + lane;
    double* Wu = Wl + n_panel * 32;
    const int nu = n_tri - n_panel;
    int e = warp;
    for (; e + 3 * NW < nu; e += 4 * NW) {
      double v[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) v[q] = Wu[(e + q * NW) * 32];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        Ug[int64_t(e + q * NW) * 32] = v[q];
        Wu[(e + q * NW) * 32] = 0.0;
      }
    }
    for (; e < nu; e += NW) {
      Ug[int64_t(e) * 32] = Wu[e * 32];
      Wu[e * 32] = 0.0;
    }
  }
  if (warp == 0) {
    stats_out[0] = pos;
    stats_out[1] = neg;
    stats_out[2] = zero;
    stats_out[3] = zpiv;
    const unsigned long long bits = __double_as_longlong(min_abs);
    stats_out[4] = static_cast<int>(bits & 0xffffffffull);
    stats_out[5] = static_cast<int>(bits >> 32);
  }
}

/// stats: per instance 8 ints (n_pos n_neg n_zero zero_pivot | min|D| bits | pad).
template <bool kFused>
__global__ void __launch_bounds__(kBatchWarps * 32, SLPB_BATCH_MIN_BLOCKS)
k_batch_factor(BatchView T, const double* __restrict__ Kb,
               const double* __restrict__ delta,
               const double* __restrict__ gamma, double* __restrict__ Pb,
               double* Ub, double* __restrict__ Db, int32_t* __restrict__ stats,
               double* gscratch, int lbuf_offset_doubles) {
  extern __shared__ double smem[];
  __shared__ int s_ticket;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total = T.n_super * T.groups;
  int32_t* fcount = T.sync + 4;
  // invariant: the frontal workspace is all zeros between tasks
  for (int e = threadIdx.x; e < T.tri_cap * 32; e += blockDim.x) smem[e] = 0.0;
  int next = 0;
  if (threadIdx.x == 0) next = atomicAdd(&T.sync[0], 1);
  for (;;) {
    if (threadIdx.x == 0) s_ticket = next;
    __syncthreads();
    const int t = s_ticket;
    if (t >= total) break;
    // the next ticket is drawn now: its latency hides behind this task
    if (threadIdx.x == 0) next = atomicAdd(&T.sync[0], 1);
    // the groups of one front hold neighbouring tickets (metadata stays hot);
    // children of (s, g) are (c, g) with smaller tickets
    const int s = T.order[t / T.groups];
    const int g = t % T.groups;
    const FrontMeta fm = load_front_meta(T.metas + s);
    const int n_tri = fm.F * (fm.F + 1) / 2;
    int st[6];
    const int inst = g * 32 + lane;
    const int* dep = &fcount[size_t(g) * T.n_super + s];
    if (n_tri <= T.tri_cap) {
      batch_factor_front<true, kFused>(warp, lane, fm, T, g, Kb, delta[inst],
                               gamma[inst], Pb, Ub, Db, smem,
                               smem + lbuf_offset_doubles, dep, st);
    } else {
      // rare oversized front: global scratch of this block (L2-resident,
      // zero-initialised at creation and left zero by every task)
      const size_t per_block = (size_t(64) * 65 / 2 + kBatchRank * 64) * 32;
      double* W = gscratch + size_t(blockIdx.x) * per_block;
      batch_factor_front<false, kFused>(warp, lane, fm, T, g, Kb, delta[inst],
                                gamma[inst], Pb, Ub, Db, W,
                                W + size_t(64) * 65 / 2 * 32, dep, st);
    }
    __syncthreads();  // also protects s_ticket and W against the next task
    if (threadIdx.x == 0 && fm.parent >= 0) {
      red_release_add(&fcount[size_t(g) * T.n_super + fm.parent], 1);
    }
    if (warp == 0) {
      int32_t* vs = stats + size_t(inst) * 8;
      if (st[0]) atomicAdd(&vs[0], st[0]);
      if (st[1]) atomicAdd(&vs[1], st[1]);
      if (st[2]) atomicAdd(&vs[2], st[2]);
      if (st[3]) atomicOr(&vs[3], st[3]);
      const unsigned long long bits =
          (unsigned long long)(unsigned)st[4] |
          ((unsigned long long)(unsigned)st[5] << 32);
      atomicMin(reinterpret_cast<unsigned long long*>(&vs[4]), bits);
    }
  }
}

/// Forward then backward substitution of every instance; tickets [0, ns·G) are
/// forward tasks, [ns·G, 2·ns·G) backward tasks (the arithmetic of
/// ldlt_forward_front / ldlt_backward_front per lane). The panels stream
/// through once per direction in 256-byte lines.
__global__ void __launch_bounds__(128)
k_batch_solve(BatchView T, const double* __restrict__ Pb,
              const double* __restrict__ Db, const double* __restrict__ rhs,
              double* xperm, double* uvecs, double* __restrict__ sol,
              int fmax) {
  extern __shared__ double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* wl = smem + size_t(warp) * fmax * 32 + lane;
  const int ns = T.n_super, G = T.groups;
  const int total = ns * G;
  int32_t* fcount = T.sync + 4;
  int32_t* fflag = fcount + size_t(total);
  int32_t* bflag = fflag + size_t(total);
  for (;;) {
    int t = 0;
    if (lane == 0) t = atomicAdd(&T.sync[1], 1);
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t >= 2 * total) break;
    const bool fwd = t < total;
    const int tt = fwd ? t : 2 * total - 1 - t;
    const int s = T.order[tt / G];
    const int g = fwd ? tt % G : G - 1 - tt % G;
    const FrontMeta fm = load_front_meta(T.metas + s);
    const int F = fm.F, np = fm.np, c0 = fm.c0;
    const double* P = Pb + (int64_t(g) * T.panel_total + fm.panel_off) * 32 + lane;
    double* xp = xperm + int64_t(g) * T.dim * 32 + lane;
    double* uv = uvecs + int64_t(g) * T.rel_total * 32 + lane;
    const size_t fi = size_t(g) * ns + s;
    if (fwd) {
      const double* r = rhs + int64_t(g) * T.dim * 32 + lane;
      for (int i = 0; i < F; ++i) {
        wl[i * 32] = i < np ? __ldg(r + int64_t(T.perm[c0 + i]) * 32) : 0.0;
      }
      wait_children_warp(lane, &fcount[fi], fm.n_child);
      __syncwarp();
      for (int ck = 0; ck < fm.n_child; ++ck) {
        const FrontMeta cm =
            load_front_meta(T.metas + T.child_idx[fm.child_begin + ck]);
        const int mc = cm.F - cm.np;
        const int32_t* rel = T.rel_idx + cm.rel_off;
        const double* u = uv + int64_t(cm.rel_off) * 32;
        for (int i = 0; i < mc; ++i) wl[rel[i] * 32] += __ldcg(u + int64_t(i) * 32);
      }
      for (int k = 0; k < np; ++k) {
        const double yk = wl[k * 32];
        const double* Pk = P + int64_t(tri_col(k, F)) * 32;
        int i = k + 1;
        for (; i + 4 <= F; i += 4) {
          double l[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) l[q] = __ldcs(Pk + int64_t(i + q) * 32);
#pragma unroll
          for (int q = 0; q < 4; ++q) wl[(i + q) * 32] -= l[q] * yk;
        }
        for (; i < F; ++i) wl[i * 32] -= __ldcs(Pk + int64_t(i) * 32) * yk;
      }
      for (int i = 0; i < np; ++i) xp[int64_t(c0 + i) * 32] = wl[i * 32];
      for (int i = np; i < F; ++i) {
        uv[int64_t(fm.rel_off + i - np) * 32] = wl[i * 32];
      }
      __syncwarp();
      if (lane == 0) {
        if (fm.parent >= 0) red_release_add(&fcount[size_t(g) * ns + fm.parent], 1);
        st_release(&fflag[fi], 1);
      }
    } else {
      const double* Dg = Db + int64_t(g) * T.dim * 32 + lane;
      const int32_t* rows = T.rows_idx + fm.rows_off;
      wait_children_warp(lane,
                         fm.parent >= 0 ? &bflag[size_t(g) * ns + fm.parent]
                                        : &fflag[fi], 1);
      __syncwarp();
      for (int i = 0; i < F; ++i) {
        wl[i * 32] = i < np ? __ldcg(xp + int64_t(c0 + i) * 32) /
                                  __ldg(Dg + int64_t(c0 + i) * 32)
                            : __ldcg(xp + int64_t(rows[i]) * 32);
      }
      // t_k = z_k − Σ_{i ≥ np} L(i,k)·x_i, i ascending
      for (int k = 0; k < np; ++k) {
        const double* Pk = P + int64_t(tri_col(k, F)) * 32;
        double acc = wl[k * 32];
        int i = np;
        for (; i + 4 <= F; i += 4) {
          double l[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) l[q] = __ldcs(Pk + int64_t(i + q) * 32);
#pragma unroll
          for (int q = 0; q < 4; ++q) acc -= l[q] * wl[(i + q) * 32];
        }
        for (; i < F; ++i) acc -= __ldcs(Pk + int64_t(i) * 32) * wl[i * 32];
        wl[k * 32] = acc;
      }
      // L11ᵀ x = t, column-oriented
      for (int i = np - 1; i >= 1; --i) {
        const double xi = wl[i * 32];
        for (int k = 0; k < i; ++k) {
          wl[k * 32] -= __ldcs(P + int64_t(tri_col(k, F) + i) * 32) * xi;
        }
      }
      double* so = sol + int64_t(g) * T.dim * 32 + lane;
      for (int i = 0; i < np; ++i) {
        const double v = wl[i * 32];
        xp[int64_t(c0 + i) * 32] = v;
        so[int64_t(T.perm[c0 + i]) * 32] = v;
      }
      __syncwarp();
      if (lane == 0) st_release(&bflag[fi], 1);
    }
  }
}

/// dst[(g·E + e)·32 + ℓ] ← src[e] for one instance (g, ℓ): setup only.
__global__ void k_batch_scatter(const double* __restrict__ src, int64_t n,
                                double* __restrict__ dst_lane) {
  const int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e < n) dst_lane[e * 32] = src[e];
}
__global__ void k_batch_gather(const double* __restrict__ src_lane, int64_t n,
                               double* __restrict__ dst) {
  const int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e < n) dst[e] = src_lane[e * 32];
}

}  // namespace slpb

struct slpb_batch {
  // Self-contained: own stream and own copies of the symbolic arrays, so that a
  // batch outlives the solver it was created from (slpb_group).
  int device = 0;
  cudaStream_t stream = nullptr;
  slpb_counters* counters = nullptr;  // optional: where launches are counted
  int n_super = 0, dim = 0;
  slpb::DevBuf<int32_t> order, child_idx, rel_idx, rows_idx, asm_src, perm;
  slpb::DevBuf<uint8_t> col_is_primal;
  int batch = 0, groups = 0;
  int tri_cap = 0, fmax = 0, factor_warps = 1;
  int factor_blocks = 0, solve_blocks = 0, factor_smem = 0, solve_smem = 0;
  int64_t nK = 0, panel_total = 0, update_total = 0, rel_total = 0;
  slpb::DevBuf<slpb::FrontMeta> metas;
  slpb::DevBuf<int32_t> asm_tri, umap, sync, stats;
  slpb::DevBuf<double> Kb, rhs, Pb, Ub, Db, xperm, uvecs, sol, delta, gamma,
      staging, gscratch;
  cudaEvent_t ev[4] = {};
  float factor_ms = 0.0f, solve_ms = 0.0f;
  bool fused_arith = false;  // arithmetic mode of the solver it was created from
  std::string error;
};

namespace slpb {

#define CUB(call)                                                         \
  do {                                                                    \
    cudaError_t e_ = (call);                                              \
    if (e_ != cudaSuccess) {                                              \
      B->error = std::string(#call) + ": " + cudaGetErrorString(e_);      \
      return SLPB_ERR_CUDA;                                               \
    }                                                                     \
  } while (0)

inline BatchView batch_view(slpb_batch* B) {
  BatchView T{};
  T.order = B->order.p;
  T.metas = B->metas.p;
  T.child_idx = B->child_idx.p;
  T.rel_idx = B->rel_idx.p;
  T.rows_idx = B->rows_idx.p;
  T.asm_src = B->asm_src.p;
  T.asm_tri = B->asm_tri.p;
  T.umap = B->umap.p;
  T.col_is_primal = B->col_is_primal.p;
  T.perm = B->perm.p;
  T.sync = B->sync.p;
  T.n_super = B->n_super;
  T.groups = B->groups;
  T.dim = B->dim;
  T.tri_cap = B->tri_cap;
  T.nK = B->nK;
  T.panel_total = B->panel_total;
  T.update_total = B->update_total;
  T.rel_total = B->rel_total;
  return T;
}

}  // namespace slpb

extern "C" {

int slpb_batch_create(slpb_solver* S, int32_t batch, slpb_batch** out) {
  using namespace slpb;
  if (!S || !out || batch < 1) return SLPB_ERR_ARGUMENT;
  *out = nullptr;
  if (!S->analyzed) {
    return fail(S, SLPB_ERR_STATE, "slpb_batch_create needs slpb_analyze first");
  }
  const Symbolic& Y = S->sym;
  if (Y.max_front > 64) {
    return fail(S, SLPB_ERR_UNSUPPORTED,
                "batched factorisation: fronts above order 64 are not supported");
  }
  CU(cudaSetDevice(S->device));
  auto B = std::make_unique<slpb_batch>();
  B->device = S->device;
  CU(cudaStreamCreateWithFlags(&B->stream, cudaStreamNonBlocking));
  cudaStream_t bstream = B->stream;
  const AllocScope alloc_scope{bstream};
  // (on any failure below the unique_ptr frees the buffers on this stream)
  B->counters = &S->counters;
  B->n_super = Y.n_super;
  B->dim = Y.dim;
  CU(B->order.upload(Y.level_supers, bstream));
  CU(B->child_idx.upload(Y.child_idx, bstream));
  CU(B->rel_idx.upload(Y.rel_idx, bstream));
  CU(B->rows_idx.upload(Y.rows_idx, bstream));
  CU(B->asm_src.upload(Y.asm_src, bstream));
  CU(B->perm.upload(Y.perm, bstream));
  CU(B->col_is_primal.upload(Y.col_is_primal, bstream));
  B->batch = batch;
  B->groups = (batch + 31) / 32;
  const int ns = Y.n_super;
  const int64_t G = B->groups;
  // packed offsets, packed assembly positions, child → parent maps
  std::vector<FrontMeta> metas(ns);
  std::vector<int64_t> uoff(ns + 1, 0), poff(ns + 1, 0);
  for (int s = 0; s < ns; ++s) {
    const int64_t F = Y.front_dim[s];
    const int64_t np = Y.super_first[s + 1] - Y.super_first[s];
    const int64_t m = F - np;
    poff[s + 1] = poff[s] + np * F - np * (np - 1) / 2;
    uoff[s + 1] = uoff[s] + m * (m + 1) / 2;
  }
  std::vector<int32_t> nchild(ns, 0);
  for (int s = 0; s < ns; ++s) {
    nchild[s] = static_cast<int32_t>(Y.child_ptr[s + 1] - Y.child_ptr[s]);
  }
  auto tri = [](int64_t i, int64_t j, int64_t F) {
    return static_cast<int32_t>(j * F - j * (j - 1) / 2 + (i - j));
  };
  // fronts whose triangle fits the shared-memory cap take the fast path: the
  // cap is the largest order that still lets two warps share an SM, or the
  // 98th percentile of the front orders when that is smaller
  std::vector<int32_t> orders(Y.front_dim);
  std::sort(orders.begin(), orders.end());
  int f_cap = orders[std::min<size_t>(orders.size() - 1,
                                       (orders.size() * 98) / 100)];
  f_cap = std::max(f_cap, 8);
  if (const char* e = std::getenv("SLPB_BATCH_FCAP")) f_cap = std::atoi(e);
  f_cap = std::min(f_cap, 28);  // (28·29/2 + 4·28)·256 B = 132 KB per block
  f_cap = std::min<int>(f_cap, Y.max_front);
  B->tri_cap = f_cap * (f_cap + 1) / 2;
  B->fmax = Y.max_front;
  for (int s = 0; s < ns; ++s) {
    FrontMeta& fm = metas[s];
    fm.F = Y.front_dim[s];
    fm.c0 = Y.super_first[s];
    fm.np = Y.super_first[s + 1] - Y.super_first[s];
    fm.n_child = nchild[s];
    fm.child_begin = static_cast<int32_t>(Y.child_ptr[s]);
    fm.asm_begin = static_cast<int32_t>(Y.asm_ptr[s]);
    fm.asm_end = static_cast<int32_t>(Y.asm_ptr[s + 1]);
    fm.rel_off = static_cast<int32_t>(Y.rel_ptr[s]);
    fm.panel_off = poff[s];
    fm.update_off = uoff[s];
    fm.rows_off = Y.rows_ptr[s];
    fm.parent = Y.super_parent[s];
    fm.pad = 0;
    fm.ext_begin = fm.ext_chunks = fm.pad2 = fm.pad3 = 0;
  }
  // packed position of every own KKT entry; the own diagonal entries (always
  // present in the pattern) carry which regularisation they take
  std::vector<int32_t> asm_tri(Y.asm_dst.size());
  for (int s = 0; s < ns; ++s) {
    const int64_t F = Y.front_dim[s];
    const int64_t np = Y.super_first[s + 1] - Y.super_first[s];
    for (int64_t k = Y.asm_ptr[s]; k < Y.asm_ptr[s + 1]; ++k) {
      const int64_t lr = Y.asm_dst[k] % F, lc = Y.asm_dst[k] / F;
      int32_t v = tri(lr, lc, F);
      if (lr == lc && lc < np) {
        v |= Y.col_is_primal[Y.super_first[s] + lc] ? kAsmPrimalDiag : kAsmDualDiag;
      }
      asm_tri[k] = v;
    }
  }
  std::vector<int32_t> umap(std::max<int64_t>(uoff[ns], 1), 0);
  for (int c = 0; c < ns; ++c) {
    const int p = Y.super_parent[c];
    if (p < 0) continue;
    const int64_t Fp = Y.front_dim[p];
    const int64_t mc = Y.front_dim[c] - (Y.super_first[c + 1] - Y.super_first[c]);
    const int32_t* rel = Y.rel_idx.data() + Y.rel_ptr[c];
    int64_t e = uoff[c];
    for (int64_t j = 0; j < mc; ++j) {
      for (int64_t i = j; i < mc; ++i) umap[e++] = tri(rel[i], rel[j], Fp);
    }
  }
  B->nK = S->recipe.K.nnz();
  B->panel_total = poff[ns];
  B->update_total = std::max<int64_t>(uoff[ns], 1);
  B->rel_total = std::max<int64_t>(Y.rel_ptr.back(), 1);
  CU(B->metas.upload(metas, bstream));
  CU(B->asm_tri.upload(asm_tri, bstream));
  CU(B->umap.upload(umap, bstream));
  CU(B->sync.alloc(4 + 3 * size_t(G) * ns));
  CU(B->stats.alloc(size_t(G) * 32 * 8));
  CU(B->Kb.alloc(size_t(G) * B->nK * 32));
  CU(B->rhs.alloc(size_t(G) * Y.dim * 32));
  CU(B->Pb.alloc(size_t(G) * B->panel_total * 32));
  CU(B->Ub.alloc(size_t(G) * B->update_total * 32));
  CU(B->Db.alloc(size_t(G) * Y.dim * 32));
  CU(B->xperm.alloc(size_t(G) * Y.dim * 32));
  CU(B->uvecs.alloc(size_t(G) * B->rel_total * 32));
  CU(B->sol.alloc(size_t(G) * Y.dim * 32));
  CU(B->delta.alloc(size_t(G) * 32));
  CU(B->gamma.alloc(size_t(G) * 32));
  CU(B->staging.alloc(std::max<size_t>(B->nK, Y.dim)));
  CU(B->Kb.zero(bstream));
  CU(B->rhs.zero(bstream));
  CU(B->delta.zero(bstream));
  CU(B->gamma.zero(bstream));
  // launch geometry
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, S->device);
  B->factor_smem = (B->tri_cap + kBatchRank * f_cap) * 32 * 8;
  CU(raise_dynamic_smem(k_batch_factor<false>, B->factor_smem));
  CU(raise_dynamic_smem(k_batch_factor<true>, B->factor_smem));
  B->fused_arith = S->factor_arith == SLPB_ARITH_TENSOR;
  int per_sm = 1;
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
      &per_sm, k_batch_factor<false>, kBatchWarps * 32, B->factor_smem));
  per_sm = std::max(per_sm, 1);
  const int64_t tasks = int64_t(ns) * G;
  B->factor_blocks =
      static_cast<int>(std::min<int64_t>(tasks, int64_t(sms) * per_sm));
  B->solve_smem = 4 * B->fmax * 32 * 8;
  CU(raise_dynamic_smem(k_batch_solve, B->solve_smem));
  int per_sm_solve = 1;
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_solve, k_batch_solve,
                                                   128, B->solve_smem));
  per_sm_solve = std::max(per_sm_solve, 1);
  B->solve_blocks = static_cast<int>(
      std::min<int64_t>((2 * tasks + 3) / 4, int64_t(sms) * per_sm_solve));
  if (Y.max_front > f_cap) {
    const size_t per_block = (size_t(64) * 65 / 2 + kBatchRank * 64) * 32;
    CU(B->gscratch.alloc(size_t(B->factor_blocks) * per_block));
    CU(B->gscratch.zero(bstream));
  }
  for (auto& e : B->ev) CU(cudaEventCreate(&e));
  CU(cudaStreamSynchronize(bstream));
  *out = B.release();
  return SLPB_OK;
}

void slpb_batch_destroy(slpb_batch* B) {
  if (!B) return;
  cudaSetDevice(B->device);
  cudaStream_t st = B->stream;
  if (st) cudaStreamSynchronize(st);
  for (auto& e : B->ev) {
    if (e) cudaEventDestroy(e);
  }
  {
    const slpb::AllocScope alloc_scope{st};
    delete B;
  }
  if (st) cudaStreamDestroy(st);
}

int slpb_batch_size(const slpb_batch* B, int32_t* batch, int32_t* groups) {
  if (!B) return SLPB_ERR_ARGUMENT;
  if (batch) *batch = B->batch;
  if (groups) *groups = B->groups;
  return SLPB_OK;
}

int slpb_batch_set_system(slpb_batch* B, int32_t instance,
                          const double* kkt_val, const double* rhs) {
  using namespace slpb;
  if (!B || instance < 0 || instance >= B->batch) return SLPB_ERR_ARGUMENT;
  CUB(cudaSetDevice(B->device));
  const int g = instance / 32, l = instance % 32;
  const int dim = B->dim;
  if (kkt_val) {
    CUB(cudaMemcpyAsync(B->staging.p, kkt_val, B->nK * 8, cudaMemcpyHostToDevice,
                        B->stream));
    k_batch_scatter<<<blocks_for(B->nK, 256), 256, 0, B->stream>>>(
        B->staging.p, B->nK, B->Kb.p + int64_t(g) * B->nK * 32 + l);
    CUB(cudaStreamSynchronize(B->stream));
  }
  if (rhs) {
    CUB(cudaMemcpyAsync(B->staging.p, rhs, size_t(dim) * 8,
                        cudaMemcpyHostToDevice, B->stream));
    k_batch_scatter<<<blocks_for(dim, 256), 256, 0, B->stream>>>(
        B->staging.p, dim, B->rhs.p + int64_t(g) * dim * 32 + l);
    CUB(cudaStreamSynchronize(B->stream));
  }
  CUB(cudaGetLastError());
  return SLPB_OK;
}

int slpb_batch_capture(slpb_batch* B, int32_t instance, slpb_solver* src) {
  using namespace slpb;
  if (!B || !src || instance < 0 || instance >= B->batch) return SLPB_ERR_ARGUMENT;
  if (!src->finalized || src->recipe.K.nnz() != B->nK || src->dim != B->dim ||
      src->device != B->device) {
    B->error = "slpb_batch_capture: the source solver has another KKT pattern";
    return SLPB_ERR_ARGUMENT;
  }
  CUB(cudaSetDevice(B->device));
  CUB(cudaStreamSynchronize(src->stream));
  const int g = instance / 32, l = instance % 32;
  const int dim = B->dim;
  k_batch_scatter<<<blocks_for(B->nK, 256), 256, 0, B->stream>>>(
      src->Kval.p, B->nK, B->Kb.p + int64_t(g) * B->nK * 32 + l);
  k_batch_scatter<<<blocks_for(dim, 256), 256, 0, B->stream>>>(
      src->rhs.p, dim, B->rhs.p + int64_t(g) * dim * 32 + l);
  CUB(cudaGetLastError());
  CUB(cudaStreamSynchronize(B->stream));
  return SLPB_OK;
}

int slpb_batch_factor(slpb_batch* B, const double* delta, const double* gamma,
                      slpb_factor_info* info) {
  using namespace slpb;
  if (!B || !delta || !gamma) return SLPB_ERR_ARGUMENT;
  CUB(cudaSetDevice(B->device));
  const int ns = B->n_super;
  const size_t lanes = size_t(B->groups) * 32;
  std::vector<double> d(lanes, 1.0), gm(lanes, 1.0);  // padding lanes: identity-ish
  std::copy(delta, delta + B->batch, d.begin());
  std::copy(gamma, gamma + B->batch, gm.begin());
  CUB(cudaMemcpyAsync(B->delta.p, d.data(), lanes * 8, cudaMemcpyHostToDevice,
                      B->stream));
  CUB(cudaMemcpyAsync(B->gamma.p, gm.data(), lanes * 8, cudaMemcpyHostToDevice,
                      B->stream));
  std::vector<int32_t> init(lanes * 8, 0);
  {
    const double inf = INFINITY;
    for (size_t i = 0; i < lanes; ++i) std::memcpy(&init[i * 8 + 4], &inf, 8);
  }
  CUB(cudaMemcpyAsync(B->stats.p, init.data(), init.size() * 4,
                      cudaMemcpyHostToDevice, B->stream));
  CUB(cudaMemsetAsync(B->sync.p, 0, (4 + size_t(B->groups) * ns) * 4, B->stream));
  CUB(cudaEventRecord(B->ev[0], B->stream));
  const BatchView T = batch_view(B);
  if (B->fused_arith) {
    k_batch_factor<true><<<B->factor_blocks, kBatchWarps * 32, B->factor_smem,
                           B->stream>>>(T, B->Kb.p, B->delta.p, B->gamma.p,
                                        B->Pb.p, B->Ub.p, B->Db.p, B->stats.p,
                                        B->gscratch.p, B->tri_cap * 32);
  } else {
    k_batch_factor<false><<<B->factor_blocks, kBatchWarps * 32, B->factor_smem,
                            B->stream>>>(T, B->Kb.p, B->delta.p, B->gamma.p,
                                         B->Pb.p, B->Ub.p, B->Db.p, B->stats.p,
                                         B->gscratch.p, B->tri_cap * 32);
  }
  CUB(cudaEventRecord(B->ev[1], B->stream));
  CUB(cudaGetLastError());
  std::vector<int32_t> host(lanes * 8);
  CUB(cudaMemcpyAsync(host.data(), B->stats.p, host.size() * 4,
                      cudaMemcpyDeviceToHost, B->stream));
  CUB(cudaStreamSynchronize(B->stream));
  CUB(cudaEventElapsedTime(&B->factor_ms, B->ev[0], B->ev[1]));
  if (B->counters) {
    ++B->counters->kernel_launches;
    B->counters->factorizations += B->batch;
    B->counters->factorizations_completed += B->batch;
  }
  if (info) {
    for (int i = 0; i < B->batch; ++i) {
      info[i].n_pos = host[size_t(i) * 8 + 0];
      info[i].n_neg = host[size_t(i) * 8 + 1];
      info[i].n_zero = host[size_t(i) * 8 + 2];
      info[i].zero_pivot = host[size_t(i) * 8 + 3];
      std::memcpy(&info[i].min_abs_d, &host[size_t(i) * 8 + 4], 8);
    }
  }
  return SLPB_OK;
}

int slpb_batch_solve(slpb_batch* B) {
  using namespace slpb;
  if (!B) return SLPB_ERR_ARGUMENT;
  CUB(cudaSetDevice(B->device));
  const int ns = B->n_super;
  CUB(cudaMemsetAsync(B->sync.p, 0, (4 + 3 * size_t(B->groups) * ns) * 4,
                      B->stream));
  CUB(cudaEventRecord(B->ev[2], B->stream));
  const BatchView T = batch_view(B);
  k_batch_solve<<<B->solve_blocks, 128, B->solve_smem, B->stream>>>(
      T, B->Pb.p, B->Db.p, B->rhs.p, B->xperm.p, B->uvecs.p, B->sol.p, B->fmax);
  CUB(cudaEventRecord(B->ev[3], B->stream));
  CUB(cudaGetLastError());
  CUB(cudaStreamSynchronize(B->stream));
  CUB(cudaEventElapsedTime(&B->solve_ms, B->ev[2], B->ev[3]));
  if (B->counters) {
    ++B->counters->kernel_launches;
    B->counters->solves += B->batch;
  }
  return SLPB_OK;
}

int slpb_batch_get(slpb_batch* B, int32_t instance, int what, double* dst) {
  using namespace slpb;
  if (!B || !dst || instance < 0 || instance >= B->batch) return SLPB_ERR_ARGUMENT;
  CUB(cudaSetDevice(B->device));
  const int g = instance / 32, l = instance % 32;
  const int dim = B->dim;
  const double* src = nullptr;
  switch (what) {
    case SLPB_BATCH_SOLUTION: src = B->sol.p; break;
    case SLPB_BATCH_D: src = B->Db.p; break;
    default: return SLPB_ERR_ARGUMENT;
  }
  k_batch_gather<<<blocks_for(dim, 256), 256, 0, B->stream>>>(
      src + int64_t(g) * dim * 32 + l, dim, B->staging.p);
  CUB(cudaGetLastError());
  CUB(cudaMemcpyAsync(dst, B->staging.p, size_t(dim) * 8, cudaMemcpyDeviceToHost,
                      B->stream));
  CUB(cudaStreamSynchronize(B->stream));
  return SLPB_OK;
}

int slpb_batch_last_ms(const slpb_batch* B, float* factor_ms, float* solve_ms) {
  if (!B) return SLPB_ERR_ARGUMENT;
  if (factor_ms) *factor_ms = B->factor_ms;
  if (solve_ms) *solve_ms = B->solve_ms;
  return SLPB_OK;
}

int slpb_batch_bytes(const slpb_batch* B, int64_t* panel_entries,
                     int64_t* update_entries) {
  if (!B) return SLPB_ERR_ARGUMENT;
  if (panel_entries) *panel_entries = B->panel_total;
  if (update_entries) *update_entries = B->update_total;
  return SLPB_OK;
}

}  // extern "C"
