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

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Exact-erf GELU (nn.GELU default) through Abramowitz-Stegun 7.1.26: |erf error| <= 1.5e-7 (fp32 level, far below the
// bf16 rounding of the stored activation), ~14 instructions with one rcp and one ex2 instead of ~25 for erff().
// tail(x) = 0.5 * erfc(|x| / sqrt2) is evaluated directly, so Phi(x) has no cancellation for very negative x.
// e = exp(-x^2 / 2) is shared with the derivative:  gelu'(x) = Phi(x) + x * e / sqrt(2 pi).
__device__ __forceinline__ float gelu_tail(float x, float& e) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
  e = ex2_approx(-1.4426950408889634f * z * z);
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  return 0.5f * p * t * e;
}
__device__ __forceinline__ float gelu_erf(float x) {
  float e;
  const float tail = gelu_tail(x, e);
  return x * (x >= 0.f ? 1.0f - tail : tail);
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
  float e;
  const float tail = gelu_tail(x, e);
  const float cdf = x >= 0.f ? 1.0f - tail : tail;
  return fmaf(x * 0.39894228040143268f, e, cdf);
}

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace eavit
