// Development aid: what programmatic dependent launch is worth for a chain of
// small dependent kernels (the shape of the solver's per-iteration sequence).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 scripts/micro/pdl_probe.cu -o scripts/micro/pdl_probe
#include <cstdio>
#include <cuda_runtime.h>

template <bool kPdl>
__global__ void k_small(const double* __restrict__ in, double* __restrict__ out, int n) {
  if (kPdl) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;");
  }
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i] * 1.0000001 + 1.0;
}

template <bool kPdl>
float run_chain(cudaStream_t st, double* a, double* b, int n, int len, bool attr, bool stream_attr_ok) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 20; ++rep) {
    cudaEventRecord(e0, st);
    for (int k = 0; k < len; ++k) {
      double* src = (k & 1) ? b : a;
      double* dst = (k & 1) ? a : b;
      if (attr) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((n + 255) / 256); cfg.blockDim = dim3(256); cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, k_small<kPdl>, (const double*)src, dst, n);
      } else {
        k_small<kPdl><<<(n + 255) / 256, 256, 0, st>>>(src, dst, n);
      }
    }
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    best = ms < best ? ms : best;
  }
  (void)stream_attr_ok;
  return best;
}

int main() {
  const int n = 45000, len = 24;
  double *a, *b;
  cudaMalloc(&a, n * 8); cudaMalloc(&b, n * 8);
  cudaMemset(a, 0, n * 8); cudaMemset(b, 0, n * 8);
  cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  const float plain = run_chain<false>(st, a, b, n, len, false, false);
  const float pdl = run_chain<true>(st, a, b, n, len, true, false);
  // does the attribute work as a stream attribute (so that <<<>>> launches inherit it)?
  cudaLaunchAttributeValue v{};
  v.programmaticStreamSerializationAllowed = 1;
  cudaError_t e = cudaStreamSetAttribute(st, cudaLaunchAttributeProgrammaticStreamSerialization, &v);
  printf("stream attribute: %s\n", cudaGetErrorString(e));
  cudaGetLastError();
  const float viastream = run_chain<true>(st, a, b, n, len, false, e == cudaSuccess);
  printf("chain of %d dependent kernels over %d doubles: plain %.1f us (%.2f per kernel), PDL attribute %.1f us (%.2f), "
         "<<<>>> after stream attribute %.1f us\n", len, n, plain * 1e3, plain * 1e3 / len, pdl * 1e3, pdl * 1e3 / len,
         viastream * 1e3);
  double h[4]; cudaMemcpy(h, a, 32, cudaMemcpyDeviceToHost);
  printf("check %.6f\n", h[0]);
  return 0;
}
