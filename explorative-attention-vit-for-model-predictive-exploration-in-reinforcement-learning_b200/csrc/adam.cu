// Fused multi-tensor Adam over one flat fp32 parameter / gradient buffer (all trainable tensors of the agent
// live in one allocation, so the optimiser is one launch and the gradient all-reduce is one NCCL call).
// Writes the fp32 master weights and the bf16 shadow used by the tensor-core GEMMs in the same pass.
//
// Reference: torch.optim.Adam(lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0) at agents.py:129 / :508.
// The step counter lives on the device so the launch can be replayed from a CUDA graph.
#include "common.cuh"

namespace eavit {

// The C ABI carries the betas as float32; the decimal constants the caller meant (0.9, 0.999: at most 7 significant
// digits survive float32 anyway) are recovered so that 1 - beta and beta ** step match torch's double arithmetic.
static inline double decimal_beta(float b) { return nearbyint((double)b * 1e7) / 1e7; }

__global__ void adam_tick_kernel(long long* step) { step[0] += 1; }

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, __nv_bfloat16* __restrict__ p_bf16, long long n4,
                                                   const long long* __restrict__ step, float lr, double beta1, double beta2,
                                                   float eps, float grad_scale) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  // torch evaluates 1 - beta, beta ** step and the bias corrections in Python doubles and hands float32 scalars to the
  // tensor ops; (float)(1 - 0.999) and 1.f - 0.999f differ by 1.3e-5 relative, which would sit in exp_avg_sq forever
  const double t = (double)step[0];
  const float bc1 = (float)(1.0 - pow(beta1, t));
  const float bc2_sqrt = (float)sqrt(1.0 - pow(beta2, t));
  const float step_size = (float)((double)lr / (1.0 - pow(beta1, t)));
  const float omb1 = (float)(1.0 - beta1), omb2 = (float)(1.0 - beta2), b2f = (float)beta2;
  (void)bc1;
  float4 pp = reinterpret_cast<float4*>(p)[i];
  float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i);
  float4 mm = reinterpret_cast<float4*>(m)[i];
  float4 vv = reinterpret_cast<float4*>(v)[i];
  float* pa = &pp.x; float* ga = &gg.x; float* ma = &mm.x; float* va = &vv.x;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float gk = ga[k] * grad_scale;
    ma[k] = ma[k] + (gk - ma[k]) * omb1;                     // exp_avg.lerp_(grad, 1 - beta1)
    va[k] = va[k] * b2f + omb2 * gk * gk;                    // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    const float denom = sqrtf(va[k]) / bc2_sqrt + eps;
    pa[k] = pa[k] - step_size * (ma[k] / denom);
  }
  reinterpret_cast<float4*>(p)[i] = pp;
  reinterpret_cast<float4*>(m)[i] = mm;
  reinterpret_cast<float4*>(v)[i] = vv;
  if (p_bf16 != nullptr)
    reinterpret_cast<uint2*>(p_bf16)[i] = make_uint2(pack_bf16x2(pp.x, pp.y), pack_bf16x2(pp.z, pp.w));
}

// sum of squares of a flat fp32 buffer -> out[0] (+=), for utils.py:141-170 global_grad_norm_ and optional clipping
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ x, long long n4, float* __restrict__ out) {
  __shared__ float red[8];
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    acc += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) { float t = 0.f; for (int i = 0; i < 8; ++i) t += red[i]; atomicAdd(out, t); }
}

// nn.utils.clip_grad_norm_: g *= min(1, max_norm / (sqrt(sumsq) + 1e-6))
__global__ void clip_scale_kernel(float* __restrict__ g, long long n4, const float* __restrict__ sumsq, float max_norm) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float coef = fminf(1.f, max_norm / (sqrtf(sumsq[0]) + 1e-6f));
  float4 v = reinterpret_cast<float4*>(g)[i];
  v.x *= coef; v.y *= coef; v.z *= coef; v.w *= coef;
  reinterpret_cast<float4*>(g)[i] = v;
}

}  // namespace eavit

using namespace eavit;

extern "C" {

int eavit_adam_step(float* p, const float* g, float* m, float* v, void* p_bf16, long long n, long long* step, float lr,
                    float beta1, float beta2, float eps, float grad_scale, void* stream) {
  EAVIT_CHECK_ARG(p && g && m && v && step && n > 0 && n % 4 == 0);
  cudaStream_t st = (cudaStream_t)stream;
  adam_tick_kernel<<<1, 1, 0, st>>>(step);
  EAVIT_LAUNCH_OK();
  adam_kernel<<<cdiv(n / 4, 256), 256, 0, st>>>(p, g, m, v, (__nv_bfloat16*)p_bf16, n / 4, step, lr, decimal_beta(beta1), decimal_beta(beta2), eps,
                                                grad_scale);
  EAVIT_LAUNCH_OK();
  return EAVIT_OK;
}

// The same update for ONE contiguous range of the flat buffers without advancing the step counter: a store with frozen
// tensors (train.py:261-263 sets requires_grad = False on model.feature.*; torch.optim.Adam then skips them) is updated
// as eavit_adam_tick + one eavit_adam_apply per trainable range.
int eavit_adam_tick(long long* step, void* stream) {
  EAVIT_CHECK_ARG(step != nullptr);
  adam_tick_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step);
  EAVIT_LAUNCH_OK();
  return EAVIT_OK;
}

int eavit_adam_apply(float* p, const float* g, float* m, float* v, void* p_bf16, long long n, const long long* step, float lr,
                     float beta1, float beta2, float eps, float grad_scale, void* stream) {
  EAVIT_CHECK_ARG(p && g && m && v && step && n > 0 && n % 4 == 0);
  adam_kernel<<<cdiv(n / 4, 256), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, (__nv_bfloat16*)p_bf16, n / 4, step, lr, decimal_beta(beta1),
                                                                  decimal_beta(beta2), eps, grad_scale);
  EAVIT_LAUNCH_OK();
  return EAVIT_OK;
}

int eavit_sumsq_f32(const float* x, long long n, float* out, void* stream) {
  EAVIT_CHECK_ARG(x && out && n > 0 && n % 4 == 0);
  int grid = cdiv(n / 4, 256);
  if (grid > 4 * kNumSMs) grid = 4 * kNumSMs;
  sumsq_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, n / 4, out);
  EAVIT_LAUNCH_OK();
  return EAVIT_OK;
}

int eavit_clip_by_norm(float* g, long long n, const float* sumsq, float max_norm, void* stream) {
  EAVIT_CHECK_ARG(g && sumsq && n > 0 && n % 4 == 0);
  clip_scale_kernel<<<cdiv(n / 4, 256), 256, 0, (cudaStream_t)stream>>>(g, n / 4, sumsq, max_norm);
  EAVIT_LAUNCH_OK();
  return EAVIT_OK;
}

}  // extern "C"
