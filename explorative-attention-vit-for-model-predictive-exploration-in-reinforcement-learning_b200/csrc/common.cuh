// Shared helpers for the eavit_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/eavit_b200.h"

namespace eavit {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define EAVIT_CHECK_ARG(cond)                                                        \
  do {                                                                               \
    if (!(cond)) {                                                                   \
      eavit::set_error("%s:%d: invalid argument: %s", __FILE__, __LINE__, #cond);    \
      return EAVIT_EINVAL;                                                           \
    }                                                                                \
  } while (0)

#define EAVIT_CUDA(expr)                                                             \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) {                                                         \
      eavit::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return EAVIT_ECUDA;                                                            \
    }                                                                                \
  } while (0)

// after a <<<>>> launch
#define EAVIT_LAUNCH_OK()                                                            \
  do {                                                                               \
    eavit::count_launch();                                                           \
    EAVIT_CUDA(cudaGetLastError());                                                  \
  } while (0)

constexpr int kNumSMs = 148;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 t = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(t);
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
  const float pdf = 0.39894228040143268f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace eavit
