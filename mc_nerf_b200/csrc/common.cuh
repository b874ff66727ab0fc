// Shared helpers for libmcnerf.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "mcnerf.h"

void mcnerf_set_error(const char* fmt, ...);
void mcnerf_count_launch(int n = 1);

#define MC_ARG(cond)                                                                      \
  do {                                                                                    \
    if (!(cond)) {                                                                        \
      mcnerf_set_error("%s:%d: bad argument: %s", __FILE__, __LINE__, #cond);             \
      return MCNERF_E_ARG;                                                                \
    }                                                                                     \
  } while (0)

#define MC_LAUNCHED()                                                                     \
  do {                                                                                    \
    mcnerf_count_launch();                                                                \
    cudaError_t e_ = cudaGetLastError();                                                  \
    if (e_ != cudaSuccess) {                                                              \
      mcnerf_set_error("%s:%d: launch failed: %s", __FILE__, __LINE__, cudaGetErrorString(e_)); \
      return (int)e_;                                                                     \
    }                                                                                     \
  } while (0)

#define MC_CUDA(call)                                                                     \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess) {                                                              \
      mcnerf_set_error("%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      return (int)e_;                                                                     \
    }                                                                                     \
  } while (0)

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float softplus_f(float x) {
  // torch.nn.Softplus(beta=1, threshold=20): x for x > 20, else log1p(exp(x))
  return x > 20.f ? x : log1pf(expf(x));
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

// torch.linspace(start, end, steps)[k] in fp32 (symmetric evaluation, as ATen does it)
__device__ __forceinline__ float linspace_f(float start, float end, int steps, int k) {
  float step = (end - start) / (float)(steps - 1);
  return (k < steps / 2) ? start + step * (float)k : end - step * (float)(steps - 1 - k);
}
