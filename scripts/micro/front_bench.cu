// Micro-benchmark (development aid, not part of the product): cycles of the
// per-front elimination with 1 warp and with a full grid.
#include <cstdio>
#include <cuda_runtime.h>
#include "ldlt_warp.cuh"
using namespace slpb;

__global__ void k(const double* Win, double* D, double* P, double* U, long long* cyc, int F, int np, int reps) {
  extern __shared__ double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* W = smem + warp * 1024;
  for (int i = lane; i < F * F; i += 32) W[i] = Win[i];
  __syncwarp();
  const int gw = blockIdx.x * (blockDim.x >> 5) + warp;
  __shared__ int ls[8][6];
  double rr = lane;
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    ldlt_eliminate_rows(lane, F, np, F - np, W, D + gw * 32, P + gw * 1024, U + gw * 1024, ls[warp], rr);
    __syncwarp();
  }
  long long t1 = clock64();
  if (lane == 0) { cyc[gw] = (t1 - t0) / reps; D[gw * 32 + 31] = ls[warp][0] + rr; }
}

int main() {
  const int F = 20, np = 11;
  double hW[1024];
  for (int j = 0; j < F; ++j) for (int i = 0; i < F; ++i) hW[i + j * F] = (i == j) ? 4.0 + i : (i > j ? 0.01 * (i + j) : 0.0);
  double *W, *D, *P, *U; long long* cyc;
  const int maxw = 296 * 8;
  cudaMalloc(&W, sizeof(hW)); cudaMemcpy(W, hW, sizeof(hW), cudaMemcpyHostToDevice);
  cudaMalloc(&D, maxw * 32 * 8); cudaMalloc(&P, maxw * 1024 * 8); cudaMalloc(&U, maxw * 1024 * 8); cudaMalloc(&cyc, maxw * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 1024 * 8);
  for (int cfg = 0; cfg < 3; ++cfg) {
    int blocks = cfg == 0 ? 1 : (cfg == 1 ? 148 : 296), threads = cfg == 0 ? 32 : 256;
    k<<<blocks, threads, (threads / 32) * 1024 * 8>>>(W, D, P, U, cyc, F, np, 20);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[8]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("blocks %d threads %d: %s cycles/front %lld %lld (F=%d np=%d => %.0f cycles/pivot)\n", blocks, threads, cudaGetErrorString(e), h[0], h[1], F, np, double(h[0]) / np);
  }
  return 0;
}
