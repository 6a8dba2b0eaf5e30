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

__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Exact-erf GELU (nn.GELU default) through Abramowitz-Stegun 7.1.26: |erf error| <= 1.5e-7 (fp32 level, far below the
// bf16 rounding of the stored activation).  The GEMM epilogues that evaluate it are instruction-issue bound (ncu: 80 %
// issue-active), so the constants are folded until 13 instructions remain (one rcp, one ex2):
//   tail(x) = 0.5 * erfc(|x| / sqrt2) = (a1' t + ... + a5' t^5) * e,  t = 1 / (1 + 0.2316419 |x|),  e = exp(-x^2 / 2)
//   gelu(x) = max(x, 0) - |x| * tail(x)          (no cancellation for very negative x, no select)
//   gelu'(x) = Phi(x) + x * e / sqrt(2 pi),      Phi(x) = [x >= 0] - copysign(tail, x)
__device__ __forceinline__ float gelu_tail(float x, float& e) {
  const float t = rcp_approx(fmaf(fabsf(x), 0.2316418882f, 1.0f));     // 0.3275911 / sqrt(2)
  const float a = x * 0.8493218003f;                                   // sqrt(log2(e) / 2): e = 2^(-a^2)
  e = ex2_approx(-(a * a));
  float p = fmaf(0.5307027145f, t, -0.7265760135f);                    // 0.5 * A&S coefficients
  p = fmaf(p, t, 0.7107068705f);
  p = fmaf(p, t, -0.142248368f);
  p = fmaf(p, t, 0.127414796f);
  return p * t * e;
}
__device__ __forceinline__ float gelu_erf(float x) {
  float e;
  const float tail = gelu_tail(x, e);
  return fmaf(-fabsf(x), tail, fmaxf(x, 0.f));
}
// gelu(x) and gelu'(x) from one tail / exp evaluation (the forward epilogue that also stores the derivative)
__device__ __forceinline__ float gelu_erf_and_grad(float x, float& grad) {
  float e;
  const float tail = gelu_tail(x, e);
  const float cdf = (x >= 0.f ? 1.0f : 0.0f) - copysignf(tail, x);
  grad = fmaf(x * 0.39894228040143268f, e, cdf);
  return fmaf(-fabsf(x), tail, fmaxf(x, 0.f));
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
  float e;
  const float tail = gelu_tail(x, e);
  const float cdf = (x >= 0.f ? 1.0f : 0.0f) - copysignf(tail, x);
  return fmaf(x * 0.39894228040143268f, e, cdf);
}

// ---------------------------------------------------------------------------------------------------- dropout
// nn.Dropout (vit.py:31,33,45,56,158; HF hidden / attention-probs dropout) as a counter-based mask: the keep decision of
// element (row r, column c) of a dropout SITE is a pure function of (seed, r, c), so the backward kernels regenerate
// the forward mask instead of storing it and every kernel layout (epilogue tiles, TMEM rows, LN warps) sees the same
// bits.  One 32-bit hash serves a column pair (16 bits each): keep iff bits >= thresh, thresh = round(p * 65536)
// (p = 0.1 -> 6554 / 65536 = 0.100006); kept values are scaled by 1 / (1 - thresh / 65536) (exactly unbiased).
// `eavit_dropout_mask` materialises the same mask for tests (the oracle is run with identical masks).
// `epoch` points at a per-device counter owned by the library (api.cu).  It is folded into every row key, so a CAPTURED
// launch (CUDA graph of the rollout forward, whose by-value seeds are frozen at capture time) still draws fresh masks on
// every replay once the graph bumps the counter (eavit_dropout_epoch_bump).  It stays 0 outside graphs: eager forward /
// backward pairs and the parity tests see the seeds they pass.
struct DropCfg {
  uint32_t thresh;      // 0 = dropout off
  uint32_t seed_lo, seed_hi;
  float scale;
  const uint32_t* epoch;   // device pointer, NULL when dropout is off
};
const uint32_t* drop_epoch_ptr();      // host: device address of the current device's epoch word (allocated on first use)
__host__ __device__ inline DropCfg make_drop(float p, unsigned long long seed) {
  DropCfg d;
  const float pc = p < 0.f ? 0.f : (p > 0.999f ? 0.999f : p);
  d.thresh = (uint32_t)(pc * 65536.0f + 0.5f);
  d.seed_lo = (uint32_t)seed;
  d.seed_hi = (uint32_t)(seed >> 32);
  d.scale = 1.0f / (1.0f - (float)d.thresh * (1.0f / 65536.0f));
  d.epoch = nullptr;
#ifndef __CUDA_ARCH__
  if (d.thresh) d.epoch = drop_epoch_ptr();
#endif
  return d;
}
__device__ __forceinline__ uint32_t lowbias32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ uint32_t drop_row_key(const DropCfg& d, uint32_t r) {
  const uint32_t e = d.epoch ? __ldg(d.epoch) : 0u;
  return lowbias32((r ^ d.seed_lo) + e * 0x85EBCA6Bu) + d.seed_hi;
}
// 32 bits for columns (2*cp, 2*cp + 1) of a row whose key is `rk`
__device__ __forceinline__ uint32_t drop_bits(uint32_t rk, uint32_t cp) { return lowbias32(rk + cp * 0x9E3779B9u); }
__device__ __forceinline__ float drop_even(const DropCfg& d, uint32_t bits) { return (bits & 0xffffu) >= d.thresh ? d.scale : 0.f; }
__device__ __forceinline__ float drop_odd(const DropCfg& d, uint32_t bits) { return (bits >> 16) >= d.thresh ? d.scale : 0.f; }
// mask factors (0 or 1/(1-p)) for 4 consecutive columns c0..c0+3 (c0 % 4 == 0) of row key rk
__device__ __forceinline__ void drop4(const DropCfg& d, uint32_t rk, uint32_t c0, float* m) {
  const uint32_t b0 = drop_bits(rk, c0 >> 1), b1 = drop_bits(rk, (c0 >> 1) + 1);
  m[0] = drop_even(d, b0); m[1] = drop_odd(d, b0); m[2] = drop_even(d, b1); m[3] = drop_odd(d, b1);
}

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace eavit
