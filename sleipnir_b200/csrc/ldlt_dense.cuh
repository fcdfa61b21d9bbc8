// Blocked front elimination of the warp-per-front kernels, with the rank-4
// Schur-complement update of dense fronts on the FP64 tensor cores (device only).
//
// A front (order F ≤ 32) sits assembled in shared memory with a fixed leading
// dimension of 32. Its own columns are eliminated in blocks of four pivots:
//   * the 4-column panel of the block is factored with the rows in registers
//     (lane = row; pivots and multipliers travel by warp shuffle);
//   * everything behind the block then takes the four pivots at once,
//       W(i, j) ← W(i, j) + Σ_q (−l_{i,k+q}) · w_{j,k+q},  q ascending,
//     in the FUSED arithmetic mode (ldlt_core.hpp) for fronts of order
//     ≥ kDenseFrontMin as 8×8×4 FP64 tensor-core products
//     (mma.sync.aligned.m8n8k4.f64, SASS DMMA) over the 8×8 tiles of the lower
//     triangle and for smaller fronts row by row with fma(); in the REFERENCE
//     mode row by row with the product and the difference rounded separately.
// The tensor core accumulates the four terms of an entry as one chain of fused
// multiply-adds in ascending q (measured bit for bit on B200 by
// scripts/micro/dmma_probe.cu: 262 144 of 262 144 entries), which is the rule
// of the fused mode in every factorisation path — so the tensor-core path, the
// row path, the generic per-front body and the batched kernels all produce the
// same bits.
//
// Reference: the dense branch of the Newton-system solve,
// solver/util/dense_regularized_ldlt.hpp:59-136 (Eigen::LDLT once the matrix is
// ≥ 25 % dense, interior_point.hpp:340-348); here the dense kernels work on the
// dense frontal matrices of the sparse factorisation (BASELINE config 3:
// "dense supernode fronts ≥ 16×16").
#pragma once

#include "ldlt_core.hpp"

namespace slpb {

#ifndef SLPB_DENSE_FRONT_MIN
#define SLPB_DENSE_FRONT_MIN 16
#endif
/// Fronts at least this big update their trailing matrix on the tensor cores.
constexpr int kDenseFrontMin = SLPB_DENSE_FRONT_MIN;
/// Leading dimension of a front in shared memory, and the columns kept. 34:
/// the 8×8 tile accesses of the tensor-core update (row g, column 2t of the
/// tile per lane) then fall on distinct banks. Tiles of the last tile row /
/// column reach up to 6 entries past the front, and a row index ≥ 34 wraps into
/// the top of the next column (an upper-triangle entry nothing reads), so 40
/// columns hold every access.
constexpr int kFrontLd = 34;
constexpr int kFrontCols = 40;
/// (−l | w) of the current pivot block: 4 per row, rows padded to 40.
constexpr int kDenseSideRows = 40;
constexpr int kDenseSideDoubles = 2 * 4 * kDenseSideRows;
/// Shared-memory doubles of one warp's front workspace: front, side buffers,
/// right-hand side.
constexpr int kFrontRhsDoubles = 32;
constexpr int kFrontSmemDoubles =
    kFrontLd * kFrontCols + kDenseSideDoubles + kFrontRhsDoubles;

/// D(8×8) = A(8×4) · B(4×8) + C(8×8), FP64, one warp.
/// Lane ℓ, g = ℓ / 4, t = ℓ % 4: a = A(g, t), b = B(t, g), c0/c1 = C(g, 2t + {0,1}).
__device__ __forceinline__ void dmma_8x8x4(double& c0, double& c1, double a,
                                           double b) {
  asm volatile(
      "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, "
      "{%0,%1};"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

/// Elimination of the own columns of a front of order F ≤ 32 assembled in W
/// (shared memory, kFrontLd × kFrontCols, lower triangle valid), then the
/// write-out of the L panel (global P: packed lower trapezoid, tri_col of
/// ldlt_core.hpp, the diagonal slot keeps d), D and the update matrix (global U:
/// packed lower triangle of order m).
/// rhs_i: this lane's entry of the front's right-hand side; it takes part in
/// the elimination (forward substitution carried along: r_i −= l_ik · r_k, the
/// arithmetic of ldlt_forward_front) and returns y (lanes < np) and the update
/// vector (lanes ≥ np). Inertia counters are meaningful in lane 0 only.
template <bool kFused>
__device__ __noinline__ void ldlt_eliminate_front(
    int lane, int F, int np, int m, double* __restrict__ W,
    double* __restrict__ side, double* __restrict__ Dk, double* __restrict__ P,
    double* __restrict__ U, int* __restrict__ local_stats, double& rhs_i) {
#ifdef SLPB_DENSE_PROFILE
  const unsigned active_at_entry = __activemask();
  unsigned lap0_ = 0, lap1_ = 0, lap2_ = 0, lap3_ = 0, lap4_ = 0, lapt_ = clock();
#define DENSE_LAP(i) { const unsigned now_ = clock(); lap##i##_ += now_ - lapt_; lapt_ = now_; }
#else
#define DENSE_LAP(i)
#endif
  double r = rhs_i;
  const int g = lane >> 2, t = lane & 3;
  double* sideL = side;                       // −l_{row, q}
  double* sideW = side + 4 * kDenseSideRows;  // unscaled w_{row, q}
  const bool tensor = kFused && F >= kDenseFrontMin;
  // rows 32 … 39 of the side buffers are only ever read as padding
  if (lane < 8) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      sideL[(32 + lane) * 4 + q] = -0.0;
      sideW[(32 + lane) * 4 + q] = 0.0;
    }
  }
  double d_own = 1.0;  // lane k keeps pivot k
  int pcol = 0;  // tri_col(k, F) of the current pivot k
#pragma unroll 1
  for (int kb = 0; kb < np; kb += 4) {
    const int nb = min(4, np - kb);
    // ---- the block's panel, rows in registers --------------------------------
    double p[4], nl[4], wu[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const double v = W[lane + (kb + q) * kFrontLd];
      p[q] = (lane < F && q < nb && kb + q <= lane) ? v : 0.0;
      nl[q] = -0.0;
      wu[q] = 0.0;
    }
    DENSE_LAP(0)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (q < nb) {
        const int k = kb + q;
        const double wk = p[q];
        // everything that travels between lanes is read first: the shuffles
        // overlap instead of queueing behind the division (one warp issues in
        // order). r_k and the rows k+1 … of column k are final at this point.
        const double d = __shfl_sync(0xffffffffu, wk, k);
        const double rk = __shfl_sync(0xffffffffu, r, k);
        double wjk[4];
#pragma unroll
        for (int q2 = q + 1; q2 < 4; ++q2) {
          wjk[q2] = __shfl_sync(0xffffffffu, wk, (kb + q2) & 31);
        }
        if (lane == k) d_own = wk;
        // l = w / d, branch-free. 0 / d: div.rn.f64 sends a zero dividend
        // through its ~90-instruction special-case subroutine, and every lane
        // outside the column (and every structural zero inside it) would drag
        // the warp through it on every pivot. The quotient is a signed zero:
        // such lanes divide 1 / d instead (fast path) and take the zero they
        // would have got (same bits).
        const bool zero_dividend = wk == 0.0 && d == d && d != 0.0;
        double dividend;  // (opaque to the compiler, which otherwise folds the
                          // substitution away and divides the zero after all)
        asm("{ .reg .pred p; setp.ne.s32 p, %3, 0; selp.f64 %0, %1, %2, p; }"
            : "=d"(dividend)
            : "d"(1.0), "d"(wk), "r"(static_cast<int>(zero_dividend)));
        const double quot = dividend / d;
        const double signed_zero = __hiloint2double(
            (__double2hiint(wk) ^ __double2hiint(d)) & 0x80000000, 0);
        const double l = zero_dividend ? signed_zero : quot;
        if (lane >= k && lane < F) P[pcol + lane] = lane > k ? l : wk;
        pcol += F - k - 1;
        // (lanes ≥ F hold zeros in p[] and take l = ±0 through everything below;
        // rows above the pivot are masked because their l is not a multiplier)
        if (lane > k) r = r - l * rk;
        const double ml = -l;
#pragma unroll
        for (int q2 = q + 1; q2 < 4; ++q2) {
          if (kFused) {
            if (lane >= kb + q2) p[q2] = fma(ml, wjk[q2], p[q2]);
          } else {
            if (lane >= kb + q2) p[q2] = p[q2] - l * wjk[q2];
          }
        }
        nl[q] = ml;
        wu[q] = wk;
      }
    }
    DENSE_LAP(1)
    const int r0 = kb + nb;  // first row / column behind the block
    if (r0 >= F) break;
    {
      const bool live = lane >= r0 && lane < F;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        nl[q] = live ? nl[q] : -0.0;
        wu[q] = live ? wu[q] : 0.0;
      }
      double2* sl = reinterpret_cast<double2*>(sideL + lane * 4);
      double2* sw = reinterpret_cast<double2*>(sideW + lane * 4);
      sl[0] = make_double2(nl[0], nl[1]);
      sl[1] = make_double2(nl[2], nl[3]);
      sw[0] = make_double2(wu[0], wu[1]);
      sw[1] = make_double2(wu[2], wu[3]);
    }
    __syncwarp();
    DENSE_LAP(2)
    if (tensor) {
      // ---- rank-4 update of the trailing lower triangle, 8×8 tiles -----------
      // F − r0 ≤ 31: at most 4 tile rows, 10 lower tiles. All of them are
      // loaded, then multiplied, then stored, so that their latencies overlap;
      // every address is one base register plus an immediate.
      const int nt = (F - r0 + 7) >> 3;
      const double* sa = sideL + (r0 + g) * 4 + t;
      const double* sb = sideW + (r0 + g) * 4 + t;
      double* Wt = W + (r0 + g) + (r0 + 2 * t) * kFrontLd;
      double a[4], b[4], c[10][2];
#pragma unroll
      for (int ti = 0; ti < 4; ++ti) {
        // (tile rows up to r0 + 8·nt − 1 ≤ F + 6 exist in the side buffers)
        a[ti] = ti < nt ? sa[32 * ti] : -0.0;
        b[ti] = ti < nt ? sb[32 * ti] : 0.0;
      }
      {
        int idx = 0;
#pragma unroll
        for (int tj = 0; tj < 4; ++tj) {
#pragma unroll
          for (int ti = tj; ti < 4; ++ti, ++idx) {
            if (ti < nt) {
              c[idx][0] = Wt[8 * ti + 8 * tj * kFrontLd];
              c[idx][1] = Wt[8 * ti + (8 * tj + 1) * kFrontLd];
            }
          }
        }
      }
      {
        int idx = 0;
#pragma unroll
        for (int tj = 0; tj < 4; ++tj) {
#pragma unroll
          for (int ti = tj; ti < 4; ++ti, ++idx) {
            if (ti < nt) dmma_8x8x4(c[idx][0], c[idx][1], a[ti], b[tj]);
          }
        }
      }
      {
        int idx = 0;
#pragma unroll
        for (int tj = 0; tj < 4; ++tj) {
#pragma unroll
          for (int ti = tj; ti < 4; ++ti, ++idx) {
            if (ti < nt) {
              Wt[8 * ti + 8 * tj * kFrontLd] = c[idx][0];
              Wt[8 * ti + (8 * tj + 1) * kFrontLd] = c[idx][1];
            }
          }
        }
      }
    } else {
      // ---- lane = row: the same chain of four terms per entry, two columns
      // at a time so that two chains overlap ---------------------------------
      for (int j = r0; j < F; j += 2) {
        const double2 a01 = *reinterpret_cast<const double2*>(sideW + j * 4);
        const double2 a23 = *reinterpret_cast<const double2*>(sideW + j * 4 + 2);
        const double2 b01 = *reinterpret_cast<const double2*>(sideW + j * 4 + 4);
        const double2 b23 = *reinterpret_cast<const double2*>(sideW + j * 4 + 6);
        double ca = W[lane + j * kFrontLd];
        double cb = W[lane + (j + 1) * kFrontLd];
        if (kFused) {
          ca = fma(nl[0], a01.x, ca);
          cb = fma(nl[0], b01.x, cb);
          ca = fma(nl[1], a01.y, ca);
          cb = fma(nl[1], b01.y, cb);
          ca = fma(nl[2], a23.x, ca);
          cb = fma(nl[2], b23.x, cb);
          ca = fma(nl[3], a23.y, ca);
          cb = fma(nl[3], b23.y, cb);
        } else {
          // c − l·w with l = −nl: the product of the negated factor is the
          // negated product, exactly
          ca = ca + nl[0] * a01.x;
          cb = cb + nl[0] * b01.x;
          ca = ca + nl[1] * a01.y;
          cb = cb + nl[1] * b01.y;
          ca = ca + nl[2] * a23.x;
          cb = cb + nl[2] * b23.x;
          ca = ca + nl[3] * a23.y;
          cb = cb + nl[3] * b23.y;
        }
        if (lane >= j && lane < F) W[lane + j * kFrontLd] = ca;
        if (lane >= j + 1 && lane < F && j + 1 < F) W[lane + (j + 1) * kFrontLd] = cb;
      }
    }
    __syncwarp();
    DENSE_LAP(3)
  }
  // update matrix (packed lower triangle of order m), four columns in flight;
  // ucol = tri_col(jj, m) advances by m − jj − 1 per column
  {
    const double* Wu = W + lane + np * kFrontLd;
    double* Ul = U + (lane - np);
    int ucol = 0;
    for (int j0 = 0; j0 < m; j0 += 4) {
      double v[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) v[q] = Wu[(j0 + q) * kFrontLd];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int jj = j0 + q;
        if (jj < m && lane >= np + jj && lane < F) Ul[ucol] = v[q];
        ucol += m - jj - 1;
      }
    }
  }
  rhs_i = r;
  DENSE_LAP(4)
#ifdef SLPB_DENSE_PROFILE
  if (lane == 0 && lap0_ + lap1_ + lap2_ + lap3_ + lap4_ > 16000u) {
    const unsigned active_at_exit = __activemask();
    printf("   laps active %08x / %08x F=%d np=%d fused=%d: panel load %u, pivots %u, side %u, trailing %u, write-out %u\n", active_at_entry, active_at_exit, F, np, int(kFused), lap0_, lap1_, lap2_, lap3_, lap4_);
  }
#endif
  // D and the inertia bookkeeping: lane k holds pivot k
  const bool own = lane < np;
  if (own) Dk[lane] = d_own;
  {
    const double eps = 2.220446049250313e-16;
    const unsigned is_pos = __ballot_sync(0xffffffffu, own && d_own > eps);
    const unsigned is_neg = __ballot_sync(0xffffffffu, own && d_own < -eps);
    const unsigned is_own = __ballot_sync(0xffffffffu, own);
    const unsigned is_zp = __ballot_sync(0xffffffffu, own && d_own == 0.0);
    // min |d| over the pivots that are numbers (fmin semantics), as a bit
    // pattern: non-negative doubles order like unsigned integers
    const double ad = fabs(d_own);
    unsigned long long bits = (own && ad == ad)
                                  ? static_cast<unsigned long long>(__double_as_longlong(ad))
                                  : 0x7ff0000000000000ull;
    const unsigned hi = __reduce_min_sync(0xffffffffu, static_cast<unsigned>(bits >> 32));
    const unsigned lo = __reduce_min_sync(
        0xffffffffu, static_cast<unsigned>(bits >> 32) == hi
                         ? static_cast<unsigned>(bits & 0xffffffffull)
                         : 0xffffffffu);
    if (lane == 0) {
      local_stats[0] = __popc(is_pos);
      local_stats[1] = __popc(is_neg);
      local_stats[2] = __popc(is_own & ~is_pos & ~is_neg);
      local_stats[3] = is_zp != 0 ? 1 : 0;
      local_stats[4] = static_cast<int>(lo);
      local_stats[5] = static_cast<int>(hi);
    }
  }
}

}  // namespace slpb
