// Per-cluster autodiff interpreter: forward value sweep + pull-style reverse
// sweeps over a compiled cluster program (formats in internal.hpp).
//
// Replaces, for one cluster of rows at a time, the reference's update_values
// (autodiff/expression_graph.hpp:85-96) and append_triplets (:106-153). The
// arithmetic of every op follows SURVEY Appendix B / expression.hpp literally
// (e.g. ∂(l/r)/∂r = a * -l / (r * r), :634-636) and is compiled without FMA
// contraction, so that it rounds like the reference's x86-64 build; only the
// transcendental functions can differ (CUDA libdevice vs glibc, ≤ 1-2 ulp).
//
// The body is written once as a function of (tid, nthreads, LC): the kernel
// runs it with one thread block per task of up to LC = 32 clusters that share a
// program (lane = cluster, levels separated by __syncthreads), and tests/emu
// runs the very same code on the host, lane after lane, to check the program
// encoding and the task plan without a GPU. It is not a CPU path of the product:
// nothing in libslpb.so calls it on the host.
#pragma once

#include <cmath>
#include <cstdint>

#include "internal.hpp"

#if defined(__CUDACC__)
#define SLPB_HD __host__ __device__ __forceinline__
#else
#define SLPB_HD inline
#endif

// Load that bypasses the (non-coherent) L1: for data another SM wrote earlier
// in the SAME kernel launch (dependency-driven tree kernels).
#if defined(__CUDA_ARCH__)
#define SLPB_LDCG(ptr) __ldcg(ptr)
#else
#define SLPB_LDCG(ptr) (*(ptr))
#endif

namespace slpb {

SLPB_HD double ad_op_value(uint8_t op, double l, double r) {
  switch (op) {
    case SLPB_OP_SUB: return l - r;
    case SLPB_OP_ADD: return l + r;
    case SLPB_OP_DIV: return l / r;
    case SLPB_OP_MUL: return l * r;
    case SLPB_OP_NEG: return -l;
    case SLPB_OP_ABS: return fabs(l);
    case SLPB_OP_ACOS: return acos(l);
    case SLPB_OP_ASIN: return asin(l);
    case SLPB_OP_ATAN: return atan(l);
    case SLPB_OP_ATAN2: return atan2(l, r);
    case SLPB_OP_CBRT: return cbrt(l);
    case SLPB_OP_COS: return cos(l);
    case SLPB_OP_COSH: return cosh(l);
    case SLPB_OP_ERF: return erf(l);
    case SLPB_OP_EXP: return exp(l);
    case SLPB_OP_HYPOT: return hypot(l, r);
    case SLPB_OP_IS_NONNEG: return l >= 0.0 ? 1.0 : 0.0;
    case SLPB_OP_IS_POS: return l > 0.0 ? 1.0 : 0.0;
    case SLPB_OP_LOG: return log(l);
    case SLPB_OP_LOG10: return log10(l);
    case SLPB_OP_MAX: return l < r ? r : l;   // std::max(l, r)
    case SLPB_OP_MIN: return r < l ? r : l;   // std::min(l, r)
    case SLPB_OP_POW: return pow(l, r);
    case SLPB_OP_SIGN: return l < 0.0 ? -1.0 : (l == 0.0 ? 0.0 : 1.0);
    case SLPB_OP_SIN: return sin(l);
    case SLPB_OP_SINH: return sinh(l);
    case SLPB_OP_SQRT: return sqrt(l);
    case SLPB_OP_TAN: return tan(l);
    case SLPB_OP_TANH: return tanh(l);
    default: return 0.0;
  }
}

/// Adjoint-weighted partial of parent `op(l, r)` (adjoint a) w.r.t. its left
/// (side 0) or right (side 1) argument.
SLPB_HD double ad_op_grad(uint8_t op, uint8_t side, double a, double l,
                          double r) {
  if (side == 0) {
    switch (op) {
      case SLPB_OP_SUB: case SLPB_OP_ADD: return a;
      case SLPB_OP_DIV: return a / r;
      case SLPB_OP_MUL: return a * r;
      case SLPB_OP_NEG: return -a;
      case SLPB_OP_ABS: return l < 0.0 ? -a : (l > 0.0 ? a : 0.0);
      case SLPB_OP_ACOS: return -a / sqrt(1.0 - l * l);
      case SLPB_OP_ASIN: return a / sqrt(1.0 - l * l);
      case SLPB_OP_ATAN: return a / (1.0 + l * l);
      case SLPB_OP_ATAN2: return a * r / (l * l + r * r);
      case SLPB_OP_CBRT: { double c = cbrt(l); return a / (3.0 * c * c); }
      case SLPB_OP_COS: return a * -sin(l);
      case SLPB_OP_COSH: return a * sinh(l);
      case SLPB_OP_ERF:
        return a * (2.0 * 0.564189583547756286948079451560772586) *
               exp(-l * l);
      case SLPB_OP_EXP: return a * exp(l);
      case SLPB_OP_HYPOT: return a * l / hypot(l, r);
      case SLPB_OP_LOG: return a / l;
      case SLPB_OP_LOG10:
        return a / (2.302585092994045684017991454684364208 * l);
      case SLPB_OP_MAX: return l >= r ? a : 0.0;
      case SLPB_OP_MIN: return l <= r ? a : 0.0;
      case SLPB_OP_POW: return a * pow(l, r - 1.0) * r;
      case SLPB_OP_SIN: return a * cos(l);
      case SLPB_OP_SINH: return a * cosh(l);
      case SLPB_OP_SQRT: return a / (2.0 * sqrt(l));
      case SLPB_OP_TAN: { double c = cos(l); return a / (c * c); }
      case SLPB_OP_TANH: { double c = cosh(l); return a / (c * c); }
      default: return 0.0;
    }
  }
  switch (op) {
    case SLPB_OP_SUB: return -a;
    case SLPB_OP_ADD: return a;
    case SLPB_OP_DIV: return a * -l / (r * r);
    case SLPB_OP_MUL: return a * l;
    case SLPB_OP_ATAN2: return a * -l / (l * l + r * r);
    case SLPB_OP_HYPOT: return a * r / hypot(l, r);
    case SLPB_OP_MAX: return l >= r ? 0.0 : a;
    case SLPB_OP_MIN: return l <= r ? 0.0 : a;
    case SLPB_OP_POW: return a * pow(l, r) * log(l);
    default: return 0.0;
  }
}

/// Development aid (scripts/sweep_debug.py): slpb.cu defines SLPB_AD_STAMP to
/// record clock64() at the phase boundaries of one thread block.
#ifndef SLPB_AD_STAMP
#define SLPB_AD_STAMP(k)
#endif

struct NoSync {
  SLPB_HD void operator()() const {}
};

/// Instruction stream read straight from the program blob (host emulation,
/// and the reference point for the device's TMA-staged ring in slpb.cu).
struct DirectStream {
  const uint32_t* P;
  SLPB_HD const uint32_t* acquire(int b) const { return P + (P + P[10])[b]; }
  SLPB_HD void release(int) const {}
};

/// Runs one TASK: up to LC clusters that share ONE program, side by side.
/// Thread `tid` of `nthreads` (a multiple of LC) works for cluster c = tid % LC
/// on the items q, q+R, … of every level (q = tid / LC, R = nthreads / LC). On
/// the device LC = 32 makes every warp run one instruction for 32 clusters
/// (time steps) at once: the opcode is warp-uniform, the scratch of slot s and
/// lane c sits at scratch[s·LC + c] (conflict-free in shared memory), and the
/// block synchronises between levels. tests/emu runs the same code with
/// nthreads = LC, one lane after the other.
///
/// `H` is the program's header + tables (P[5] words; in shared memory on the
/// device); `stream` hands out the level blocks in order (acquire → process →
/// barrier → release). `scratch` holds H[0]·LC doubles; values and adjoints
/// share it (the compiler assigns physical slots by liveness). `B` is the
/// task's binding block, transposed so that lanes read consecutive words:
///   leaf_index i32[n_leaf][LC] | pad | const_val f64[n_const][LC]
///   | val_out i32[n_val_out][LC] | adj_out i32[n_adj_out][LC]
/// Lanes ≥ count carry a copy of lane 0's bindings and skip their stores.
template <int LC, typename Stream, typename Sync>
SLPB_HD void ad_run_group(int tid, int nthreads, int count,
                          const uint32_t* __restrict__ H, Stream& stream,
                          const uint32_t* __restrict__ B,
                          const double* __restrict__ leaf,
                          double* __restrict__ stage, double* scratch,
                          Sync sync) {
  const int c = tid % LC;
  const int q = tid / LC;
  const int R = nthreads / LC;
  const bool active = c < count;
  const int n_leaf = static_cast<int>(H[2]);
  const int n_const = static_cast<int>(H[3]);
  const int n_blocks = static_cast<int>(H[4]);
  const int n_val_out = static_cast<int>(H[6]);
  const int n_adj_out = static_cast<int>(H[7]);
  double* S = scratch + c;  // slot s of this lane: S[s * LC]

  const int32_t* leaf_index = reinterpret_cast<const int32_t*>(B);
  const int const_off = (n_leaf * LC + 1) & ~1;
  const double* const_val = reinterpret_cast<const double*>(B + const_off);
  const int32_t* val_out_stage =
      reinterpret_cast<const int32_t*>(B + const_off + 2 * n_const * LC);
  const int32_t* adj_out_stage = val_out_stage + n_val_out * LC;

  const uint16_t* leaf_slot = reinterpret_cast<const uint16_t*>(H + H[8]);
  const uint16_t* const_slot = reinterpret_cast<const uint16_t*>(H + H[9]);
  for (int i = q; i < n_leaf; i += R) {
    S[leaf_slot[i] * LC] = leaf[leaf_index[i * LC + c]];
  }
  for (int i = q; i < n_const; i += R) {
    S[const_slot[i] * LC] = const_val[i * LC + c];
  }
  sync();
  SLPB_AD_STAMP(2);

  for (int blk = 0; blk < n_blocks; ++blk) {
    const uint32_t* Bk = stream.acquire(blk);
    const uint32_t kind = Bk[0];
    const int n_items = static_cast<int>(Bk[1]);
    // workers' private lists: {kind, n_items, n_contrib, W | off[W + 1] | pad}
    const int Wb = static_cast<int>(Bk[3]);
    const uint32_t* off = Bk + 4;
    const uint32_t* body = Bk + 4 + ((Wb + 2) & ~1);
    if (kind == kBlockForward) {
      // ---- one super-level of the forward value sweep: every worker runs its
      // list in order (it reads back its own results in program order) --------
      const FwdInstr* fwd = reinterpret_cast<const FwdInstr*>(body);
      for (int w = q; w < Wb; w += R) {
        const int e = static_cast<int>(off[w + 1]);
        for (int i = static_cast<int>(off[w]); i < e; ++i) {
          const FwdInstr in = fwd[i];
          const double l = S[in.a * LC], r = S[in.b * LC];
          double v;
          if (in.op == SLPB_OP_MUL) {  // the common ops first
            v = l * r;
          } else if (in.op == SLPB_OP_ADD) {
            v = l + r;
          } else {
            v = ad_op_value(in.op, l, r);
          }
          S[in.dst * LC] = v;
        }
      }
    } else if (kind == kBlockReverse) {
      // ---- one super-level of the reverse sweeps: each visit pulls from its
      // parents' adjoints, in the row's parent order --------------------------
      const Visit* visit = reinterpret_cast<const Visit*>(body);
      const Contrib* contrib =
          reinterpret_cast<const Contrib*>(body + 2 * n_items);
      for (int w = q; w < Wb; w += R) {
        const int e = static_cast<int>(off[w + 1]);
        for (int i = static_cast<int>(off[w]); i < e; ++i) {
          const Visit v = visit[i];
          double a;
          if (v.n_contrib == 0) {
            a = static_cast<double>(v.seed);
          } else {
            a = 0.0;
            const Contrib* ct = contrib + v.contrib_begin;
            for (int k = 0; k < v.n_contrib; ++k) {
              const Contrib ck = ct[k];
              const double pa = S[ck.parent_adj * LC], lv = S[ck.l * LC];
              if (ck.op == kOpLinear) {
                a += pa * lv;  // adjoint × (±1 or the other factor): no decode
              } else if (ck.op == kOpLinearNeg) {
                a -= pa * lv;  // = a + pa·(−lv) bit for bit
              } else {
                a += ad_op_grad(ck.op, ck.side, pa, lv, S[ck.r * LC]);
              }
            }
          }
          S[v.adj * LC] = a;
        }
      }
    } else {
      // ---- value outputs (slots read here may be recycled afterwards) --------
      const uint16_t* val_out_slot =
          reinterpret_cast<const uint16_t*>(H + H[15]);
      for (int i = q; i < n_val_out; i += R) {
        if (active) stage[val_out_stage[i * LC + c]] = S[val_out_slot[i] * LC];
      }
    }
    sync();
    SLPB_AD_STAMP(3 + blk);
    stream.release(blk);
  }
  const uint16_t* adj_out_slot = reinterpret_cast<const uint16_t*>(H + H[16]);
  for (int i = q; i < n_adj_out; i += R) {
    if (active) stage[adj_out_stage[i * LC + c]] = S[adj_out_slot[i] * LC];
  }
  SLPB_AD_STAMP(62);
}

/// Scale reference of a gather source: −1 → 1, −2 → d_f, k ≥ 0 → d_c[k].
SLPB_HD double gather_entry(int e, const int32_t* __restrict__ ptr,
                            const int32_t* __restrict__ src_idx,
                            const int32_t* __restrict__ src_scale,
                            const double* __restrict__ stage, double d_f,
                            const double* __restrict__ d_c) {
  // Sources come in runs of one scale (the d_f·H_f run and the H_c run of a
  // Hessian entry; one run otherwise). A run is summed first and scaled once,
  // as the reference scales whole matrices after evaluating them
  // (problem.hpp:622-660: scaling.f * H_f.value() + H_c.value()).
  double acc = 0.0;
  const int b = ptr[e], en = ptr[e + 1];
  int k = b;
  while (k < en) {
    const int32_t sc = src_scale[k];
    double run = 0.0;
    const int k0 = k;
    for (; k < en && src_scale[k] == sc; ++k) {
      const int32_t raw = src_idx[k];
      double v = stage[raw & 0x7fffffff];
      if (raw < 0) v = -v;
      run = (k == k0) ? v : run + v;
    }
    if (sc == -2) {
      run = d_f * run;
    } else if (sc >= 0) {
      run = d_c[sc] * run;
    }
    acc = (k0 == b) ? run : acc + run;
  }
  return acc;
}

}  // namespace slpb
