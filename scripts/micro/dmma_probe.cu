// Development aid (not part of the product).
//  A. How does mma.sync.m8n8k4.f64 (DMMA) round? Random 8×4 · 4×8 + 8×8
//     products with wide exponent spreads are compared bit by bit with
//     candidate evaluation orders computed on the host.
//  B. Cycles per front of the blocked elimination (ldlt_eliminate_front) in its
//     two arithmetic modes, and their bits against host eliminations with
//     separately rounded / fused Schur-complement terms.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -I sleipnir_b200/csrc
//      -I include scripts/micro/dmma_probe.cu -o scripts/micro/dmma_probe
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include <cuda_runtime.h>

#ifdef PROFILE_LAPS
#define SLPB_DENSE_PROFILE 1
#endif

#include "ldlt_warp.cuh"
#include "ldlt_dense.cuh"
using namespace slpb;

__global__ void k_mma(const double* A, const double* B, const double* C, double* D, int n) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  for (int s = blockIdx.x; s < n; s += gridDim.x) {
    const double* a = A + s * 32;
    const double* b = B + s * 32;
    const double* c = C + s * 64;
    double c0 = c[g * 8 + 2 * t], c1 = c[g * 8 + 2 * t + 1];
    dmma_8x8x4(c0, c1, a[g * 4 + t], b[t * 8 + g]);
    D[s * 64 + g * 8 + 2 * t] = c0;
    D[s * 64 + g * 8 + 2 * t + 1] = c1;
  }
}

// dependent-issue latencies of the operations on the pivot chain
__global__ void k_lat(double* out, long long* cyc, double x0, double y0) {
  double x = x0, y = y0;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; ++i) {
    x = fma(x, y, y); x = fma(x, y, y); x = fma(x, y, y); x = fma(x, y, y);
  }
  long long t1 = clock64();
  double z = x0;
#pragma unroll 1
  for (int i = 0; i < 256; ++i) {
    z = __shfl_sync(0xffffffffu, z, (threadIdx.x + 1) & 31); z = __shfl_sync(0xffffffffu, z, (threadIdx.x + 1) & 31);
    z = __shfl_sync(0xffffffffu, z, (threadIdx.x + 1) & 31); z = __shfl_sync(0xffffffffu, z, (threadIdx.x + 1) & 31);
  }
  long long t2 = clock64();
  double w = y0;
#pragma unroll 1
  for (int i = 0; i < 256; ++i) {
    w = x0 / w; w = x0 / w; w = x0 / w; w = x0 / w;
  }
  long long t3 = clock64();
  double c0 = x0, c1 = y0;
#pragma unroll 1
  for (int i = 0; i < 256; ++i) {
    dmma_8x8x4(c0, c1, y0, y0); dmma_8x8x4(c0, c1, y0, y0); dmma_8x8x4(c0, c1, y0, y0); dmma_8x8x4(c0, c1, y0, y0);
  }
  long long t4 = clock64();
  double v = y0;
#pragma unroll 1
  for (int i = 0; i < 256; ++i) {
    v = __drcp_rn(v); v = __drcp_rn(v); v = __drcp_rn(v); v = __drcp_rn(v);
  }
  long long t5 = clock64();
  if (threadIdx.x == 0) {
    cyc[0] = (t1 - t0) / 1024; cyc[1] = (t2 - t1) / 1024; cyc[2] = (t3 - t2) / 1024; cyc[3] = (t4 - t3) / 1024; cyc[4] = (t5 - t4) / 1024;
  }
  out[threadIdx.x] = x + z + w + c0 + c1 + v;
}

template <bool kDense>
__global__ void k_front(const double* Win, double* D, double* P, double* U, long long* cyc,
                        int F, int np, int reps, double* rhs_out) {
  extern __shared__ double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int per_warp = kFrontSmemDoubles;
  double* W = smem + warp * per_warp;
  double* side = W + kFrontLd * kFrontCols;
  const int gw = blockIdx.x * (blockDim.x >> 5) + warp;
  __shared__ int ls[8][6];
  double rr = 0;
  long long total = 0;
  for (int r = 0; r < reps; ++r) {
    for (int j = 0; j < F; ++j) W[lane + j * kFrontLd] = lane < F ? Win[lane + j * F] : 0.0;
    __syncwarp();
    rr = 1.0 + lane;
    long long t0 = clock64();
    if (kDense) {
      ldlt_eliminate_front<true>(lane, F, np, F - np, W, side, D + gw * 32, P + gw * 1024,
                                 U + gw * 1024, ls[warp], rr);
    } else {
      ldlt_eliminate_front<false>(lane, F, np, F - np, W, side, D + gw * 32, P + gw * 1024,
                                  U + gw * 1024, ls[warp], rr);
    }
    __syncwarp();
    total += clock64() - t0;
  }
  if (lane == 0) cyc[gw] = total / reps;
  if (gw == 0) rhs_out[lane] = rr;
}

static void host_front(int F, int np, std::vector<double> W, bool fused, std::vector<double>& P,
                       std::vector<double>& U, std::vector<double>& D) {
  P.assign(F * np, 0.0);  // packed lower trapezoid (tri_col), padded
  D.assign(np, 0.0);
  const int m = F - np;
  U.assign(m * m + 1, 0.0);  // packed lower triangle, padded
  std::vector<double> l(F);
  for (int k = 0; k < np; ++k) {
    const double d = W[k + k * F];
    D[k] = d;
    for (int i = k + 1; i < F; ++i) l[i] = W[i + k * F] / d;
    for (int j = k + 1; j < F; ++j) {
      const double wjk = W[j + k * F];
      for (int i = j; i < F; ++i) {
        W[i + j * F] = fused ? std::fma(-l[i], wjk, W[i + j * F]) : W[i + j * F] - l[i] * wjk;
      }
    }
    P[tri_col(k, F) + k] = d;
    for (int i = k + 1; i < F; ++i) P[tri_col(k, F) + i] = l[i];
  }
  for (int j = 0; j < m; ++j)
    for (int i = j; i < m; ++i) U[tri_col(j, m) + i] = W[(np + i) + (np + j) * F];
}

int main() {
  // ---------------- A: rounding of DMMA ----------------
  {
    const int n = 4096;
    std::mt19937_64 rng(1);
    std::uniform_real_distribution<double> u(-1.0, 1.0);
    std::uniform_int_distribution<int> ex(-12, 12);
    std::vector<double> A(n * 32), B(n * 32), C(n * 64), D(n * 64);
    for (auto& v : A) v = std::ldexp(u(rng), ex(rng));
    for (auto& v : B) v = std::ldexp(u(rng), ex(rng));
    for (auto& v : C) v = std::ldexp(u(rng), ex(rng));
    double *dA, *dB, *dC, *dD;
    cudaMalloc(&dA, A.size() * 8); cudaMalloc(&dB, B.size() * 8);
    cudaMalloc(&dC, C.size() * 8); cudaMalloc(&dD, D.size() * 8);
    cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dC, C.data(), C.size() * 8, cudaMemcpyHostToDevice);
    k_mma<<<64, 32>>>(dA, dB, dC, dD, n);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(D.data(), dD, D.size() * 8, cudaMemcpyDeviceToHost);
    long long asc = 0, desc = 0, unf = 0, prod_first = 0, total = 0;
    for (int s = 0; s < n; ++s) {
      for (int i = 0; i < 8; ++i)
        for (int j = 0; j < 8; ++j) {
          const double* a = &A[s * 32 + i * 4];
          const double c = C[s * 64 + i * 8 + j];
          auto b = [&](int k) { return B[s * 32 + k * 8 + j]; };
          double x = c;
          for (int k = 0; k < 4; ++k) x = std::fma(a[k], b(k), x);
          double y = c;
          for (int k = 3; k >= 0; --k) y = std::fma(a[k], b(k), y);
          double z = c;
          for (int k = 0; k < 4; ++k) {
            volatile double pr = a[k] * b(k);
            z = z + pr;
          }
          // products summed first, then added to c
          double w = a[0] * b(0);
          for (int k = 1; k < 4; ++k) w = std::fma(a[k], b(k), w);
          w = w + c;
          const double dv = D[s * 64 + i * 8 + j];
          asc += memcmp(&dv, &x, 8) == 0;
          desc += memcmp(&dv, &y, 8) == 0;
          unf += memcmp(&dv, &z, 8) == 0;
          prod_first += memcmp(&dv, &w, 8) == 0;
          ++total;
        }
    }
    printf("[dmma] %s: %lld entries; equal to fma chain ascending k: %lld, descending k: %lld, "
           "unfused ascending: %lld, products first: %lld\n",
           cudaGetErrorString(e), total, asc, desc, unf, prod_first);
  }
  {
    double* out; long long* cyc; cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 8 * 8);
    k_lat<<<1, 32>>>(out, cyc, 1.0000001, 0.9999999);
    cudaDeviceSynchronize();
    long long h[5]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("[latency] dependent DFMA %lld cycles, 64-bit shuffle %lld, div.rn.f64 %lld, DMMA (accumulator chain) %lld, rcp.rn.f64 %lld\n", h[0], h[1], h[2], h[3], h[4]);
  }
  // ---------------- B: front elimination ----------------
  const int cfgs[][2] = {{20, 11}, {16, 8}, {24, 24}, {32, 16}, {21, 14}, {18, 3}, {12, 7}, {8, 4}, {32, 32}, {5, 1}};
  for (auto& cfg : cfgs) {
    const int F = cfg[0], np = cfg[1];
    std::vector<double> hW(1024, 0.0);
    std::mt19937_64 rng(F * 100 + np);
    std::uniform_real_distribution<double> u(-1.0, 1.0);
    for (int j = 0; j < F; ++j)
      for (int i = j; i < F; ++i) hW[i + j * F] = (i == j) ? (j % 3 == 2 ? -3.0 : 4.0) + u(rng) : u(rng);
    std::vector<double> Pf, Uf, Df, Pu, Uu, Du;
    host_front(F, np, hW, true, Pf, Uf, Df);
    host_front(F, np, hW, false, Pu, Uu, Du);
    double *W, *D, *P, *U, *rhs; long long* cyc;
    const int maxw = 296 * 4;
    cudaMalloc(&W, 1024 * 8); cudaMemcpy(W, hW.data(), 1024 * 8, cudaMemcpyHostToDevice);
    cudaMalloc(&D, maxw * 32 * 8); cudaMalloc(&P, maxw * 1024 * 8); cudaMalloc(&U, maxw * 1024 * 8);
    cudaMalloc(&cyc, maxw * 8); cudaMalloc(&rhs, 32 * 8);
    const int smem = 4 * kFrontSmemDoubles * 8;
    cudaFuncSetAttribute(k_front<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k_front<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int dense = 0; dense < 2; ++dense) {
      for (int cfgk = 0; cfgk < 2; ++cfgk) {
        const int blocks = cfgk == 0 ? 1 : 296, threads = cfgk == 0 ? 32 : 128;
        cudaMemset(P, 0, 1024 * 8); cudaMemset(U, 0, 1024 * 8);
        if (dense) k_front<true><<<blocks, threads, smem>>>(W, D, P, U, cyc, F, np, 20, rhs);
        else k_front<false><<<blocks, threads, smem>>>(W, D, P, U, cyc, F, np, 20, rhs);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[2]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        std::vector<double> gP(F * np), gU((F - np) * (F - np) + 1), gD(np);
        cudaMemcpy(gP.data(), P, F * np * 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(gU.data(), U, (F - np) * (F - np) * 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(gD.data(), D, np * 8, cudaMemcpyDeviceToHost);
        const auto& rP = dense ? Pf : Pu; const auto& rU = dense ? Uf : Uu; const auto& rD = dense ? Df : Du;
        int bad = 0;
        for (int k = 0; k < np; ++k) {
          bad += memcmp(&gD[k], &rD[k], 8) != 0;
          for (int i = k; i < F; ++i) {
            bad += memcmp(&gP[tri_col(k, F) + i], &rP[tri_col(k, F) + i], 8) != 0;
          }
        }
        const int m = F - np;
        for (int j = 0; j < m; ++j) {
          for (int i = j; i < m; ++i) bad += memcmp(&gU[tri_col(j, m) + i], &rU[tri_col(j, m) + i], 8) != 0;
        }
        printf("[front] F=%2d np=%2d %s blocks %3d x %d warps: %s, %lld cycles/front (%.0f per pivot), "
               "entries differing from the host %s elimination: %d\n",
               F, np, dense ? "tensor mode   " : "reference mode", blocks, threads / 32, cudaGetErrorString(e), h[0],
               double(h[0]) / np, dense ? "fused" : "unfused", bad);
      }
    }
    cudaFree(W); cudaFree(D); cudaFree(P); cudaFree(U); cudaFree(cyc); cudaFree(rhs);
  }
  return 0;
}
