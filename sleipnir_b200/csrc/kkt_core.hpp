// Element-wise bodies of the KKT assembly (interior_point.hpp:426-448 of the
// reference), shared by the CUDA kernels and the host emulation in tests/emu.
#pragma once

#include <cstdint>

#include "ad_core.hpp"  // SLPB_HD

namespace slpb {

/// One lower-triangle entry of lhs = [H + tril(A_iᵀΣA_i); A_e] (no δ, γ).
/// Term order follows Eigen's evaluation of (A_iᵀ Σ) A_i and H + (…).
SLPB_HD double kkt_entry(int e, const int32_t* __restrict__ h_idx,
                         const int32_t* __restrict__ ae_idx,
                         const int32_t* __restrict__ prod_ptr,
                         const int32_t* __restrict__ prod_a,
                         const int32_t* __restrict__ prod_b,
                         const int32_t* __restrict__ prod_row,
                         const double* __restrict__ Hv,
                         const double* __restrict__ Aev,
                         const double* __restrict__ Aiv,
                         const double* __restrict__ sigma) {
  const int ae = ae_idx[e];
  if (ae >= 0) return Aev[ae];
  const int pb = prod_ptr[e], pe = prod_ptr[e + 1];
  double prod = 0.0;
  for (int k = pb; k < pe; ++k) {
    const double t = (Aiv[prod_a[k]] * sigma[prod_row[k]]) * Aiv[prod_b[k]];
    prod = (k == pb) ? t : prod + t;
  }
  const int h = h_idx[e];
  if (h >= 0) return pe > pb ? Hv[h] + prod : Hv[h];
  return prod;  // 0.0 for a forced diagonal entry with no contribution
}

}  // namespace slpb
