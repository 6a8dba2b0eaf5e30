// RND target / predictor convolutions as implicit GEMM on the tcgen05 GEMM: these kernels are the
// (de)materialisation around it.  Activations are NHWC bf16 ([B*OH*OW, C] == the GEMM's [M, N] output),
// the im2col K order is torch's weight order (c, kh, kw) so Conv2d weights are used unpermuted.
//
// Reference: model.py:368-416 (conv 8/4, 4/2, 3/1 + LeakyReLU, Flatten, Linear), agents.py:333.
#include "common.cuh"

namespace eavit {

// col[(b,oy,ox), c*KH*KW + i*KW + j] = in[b, oy*s+i, ox*s+j, c]
template <typename InT>
__global__ void __launch_bounds__(256) im2col_kernel(const InT* __restrict__ in, const long long* __restrict__ sample_idx, int B, int H, int W, int C, int KH, int KW,
                                                     int stride, int OH, int OW, __nv_bfloat16* __restrict__ col, int split3) {
  const int K = C * KH * KW;
  const long long total = (long long)B * OH * OW * (K / 2);
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int k2 = (int)(i % (K / 2));
  const long long m = i / (K / 2);
  const int ox = (int)(m % OW), oy = (int)((m / OW) % OH);
  const long long b = m / ((long long)OW * OH);
  const long long sb = sample_idx ? sample_idx[b] : b;
  float v[2];
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int k = 2 * k2 + t;
    const int c = k / (KH * KW), r = k % (KH * KW), ki = r / KW, kj = r % KW;
    const size_t idx = (((size_t)sb * H + (oy * stride + ki)) * W + (ox * stride + kj)) * C + c;
    if constexpr (sizeof(InT) == 4) v[t] = in[idx]; else v[t] = __bfloat162float(in[idx]);
  }
  if (!split3) {
    *reinterpret_cast<uint32_t*>(col + (size_t)m * K + 2 * k2) = pack_bf16x2(v[0], v[1]);
  } else {
    // bf16x3 operand: [hi | hi | lo] so that A3 . B3^T (B3 = [hi | lo | hi]) = hi*hi + hi*lo + lo*hi
    const uint32_t hi = pack_bf16x2(v[0], v[1]);
    const float2 h = unpack_bf16x2(hi);
    const uint32_t lo = pack_bf16x2(v[0] - h.x, v[1] - h.y);
    __nv_bfloat16* row = col + (size_t)m * 3 * K;
    *reinterpret_cast<uint32_t*>(row + 2 * k2) = hi;
    *reinterpret_cast<uint32_t*>(row + K + 2 * k2) = hi;
    *reinterpret_cast<uint32_t*>(row + 2 * K + 2 * k2) = lo;
  }
}

// out[r, :] = mode 0: [hi | hi | lo] (activation), mode 1: [hi | lo | hi] (weight) of the fp32 row in[r, :K]
__global__ void __launch_bounds__(256) split3_rows_kernel(const float* __restrict__ in, long long ldi, int R, int K,
                                                          __nv_bfloat16* __restrict__ out, int mode) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)R * (K / 2)) return;
  const int k2 = (int)(i % (K / 2));
  const long long r = i / (K / 2);
  const float2 v = *reinterpret_cast<const float2*>(in + r * ldi + 2 * k2);
  const uint32_t hi = pack_bf16x2(v.x, v.y);
  const float2 h = unpack_bf16x2(hi);
  const uint32_t lo = pack_bf16x2(v.x - h.x, v.y - h.y);
  __nv_bfloat16* row = out + (size_t)r * 3 * K;
  *reinterpret_cast<uint32_t*>(row + 2 * k2) = hi;
  *reinterpret_cast<uint32_t*>(row + K + 2 * k2) = mode == 0 ? hi : lo;
  *reinterpret_cast<uint32_t*>(row + 2 * K + 2 * k2) = mode == 0 ? lo : hi;
}

__global__ void __launch_bounds__(256) nhwc_to_flat_f32_kernel(const float* __restrict__ act, int B, int HW, int C,
                                                               float* __restrict__ flat) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * HW * C) return;
  const int p = (int)(i % HW), c = (int)((i / HW) % C);
  const long long b = i / ((long long)HW * C);
  flat[i] = act[((size_t)b * HW + p) * C + c];
}

// d_in[b,y,x,c] = lrelu'(act[b,y,x,c]) * sum_{i,j} dcol[(b,(y-i)/s,(x-j)/s), c*KH*KW + i*KW + j]
__global__ void __launch_bounds__(256) col2im_lrelu_kernel(const __nv_bfloat16* __restrict__ dcol, const __nv_bfloat16* __restrict__ act,
                                                           int B, int H, int W, int C, int KH, int KW, int stride, int OH,
                                                           int OW, __nv_bfloat16* __restrict__ din) {
  const long long total = (long long)B * H * W * C;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  const int x = (int)((i / C) % W), y = (int)((i / ((long long)C * W)) % H);
  const long long b = i / ((long long)C * W * H);
  const int K = C * KH * KW;
  float acc = 0.f;
  for (int ki = 0; ki < KH; ++ki) {
    const int ty = y - ki;
    if (ty < 0 || ty % stride != 0) continue;
    const int oy = ty / stride;
    if (oy >= OH) continue;
    for (int kj = 0; kj < KW; ++kj) {
      const int tx = x - kj;
      if (tx < 0 || tx % stride != 0) continue;
      const int ox = tx / stride;
      if (ox >= OW) continue;
      acc += __bfloat162float(dcol[(((size_t)b * OH + oy) * OW + ox) * K + c * KH * KW + ki * KW + kj]);
    }
  }
  const float a = __bfloat162float(act[i]);
  din[i] = __float2bfloat16(a > 0.f ? acc : 0.01f * acc);
}

// Flatten of NCHW: flat[b, c*HW + p] = act[b, p, c]   (model.py:387 Flatten after NHWC conv output)
__global__ void __launch_bounds__(256) nhwc_to_flat_kernel(const __nv_bfloat16* __restrict__ act, int B, int HW, int C,
                                                           __nv_bfloat16* __restrict__ flat) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * HW * C) return;
  const int p = (int)(i % HW), c = (int)((i / HW) % C);
  const long long b = i / ((long long)HW * C);
  flat[i] = act[((size_t)b * HW + p) * C + c];
}
// backward: dact[b, p, c] = lrelu'(act[b,p,c]) * dflat[b, c*HW + p]
__global__ void __launch_bounds__(256) flat_to_nhwc_lrelu_kernel(const __nv_bfloat16* __restrict__ dflat, const __nv_bfloat16* __restrict__ act,
                                                                 int B, int HW, int C, __nv_bfloat16* __restrict__ dact) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * HW * C) return;
  const int c = (int)(i % C), p = (int)((i / C) % HW);
  const long long b = i / ((long long)HW * C);
  const float g = __bfloat162float(dflat[((size_t)b * C + c) * HW + p]);
  const float a = __bfloat162float(act[i]);
  dact[i] = __float2bfloat16(a > 0.f ? g : 0.01f * g);
}

}  // namespace eavit

using namespace eavit;

extern "C" {

int eavit_im2col(const void* in, int in_dtype, const long long* sample_idx, int B, int H, int W, int C, int KH, int KW, int stride, void* col,
                 int split3, void* stream) {
  EAVIT_CHECK_ARG(in && col && B > 0 && H >= KH && W >= KW && stride > 0 && (C * KH * KW) % 2 == 0);
  const int OH = (H - KH) / stride + 1, OW = (W - KW) / stride + 1;
  const long long total = (long long)B * OH * OW * (C * KH * KW / 2);
  cudaStream_t st = (cudaStream_t)stream;
  if (in_dtype == EAVIT_F32) im2col_kernel<float><<<cdiv(total, 256), 256, 0, st>>>((const float*)in, sample_idx, B, H, W, C, KH, KW, stride, OH, OW, (__nv_bfloat16*)col, split3);
  else if (in_dtype == EAVIT_BF16) im2col_kernel<__nv_bfloat16><<<cdiv(total, 256), 256, 0, st>>>((const __nv_bfloat16*)in, sample_idx, B, H, W, C, KH, KW, stride, OH, OW, (__nv_bfloat16*)col, split3);
  else { set_error("im2col: bad dtype %d", in_dtype); return EAVIT_EINVAL; }
  EAVIT_LAUNCH_OK();
  return EAVIT_OK;
}

int eavit_split3_rows(const float* in, long long ldi, int R, int K, void* out_bf16, int mode, void* stream) {
  EAVIT_CHECK_ARG(in && out_bf16 && R > 0 && K > 0 && K % 2 == 0 && ldi % 2 == 0 && (mode == 0 || mode == 1));
  split3_rows_kernel<<<cdiv((long long)R * (K / 2), 256), 256, 0, (cudaStream_t)stream>>>(in, ldi, R, K, (__nv_bfloat16*)out_bf16, mode);
  EAVIT_LAUNCH_OK();
  return EAVIT_OK;
}

int eavit_nhwc_to_flat_f32(const float* act, int B, int HW, int C, float* flat, void* stream) {
  EAVIT_CHECK_ARG(act && flat && B > 0 && HW > 0 && C > 0);
  nhwc_to_flat_f32_kernel<<<cdiv((long long)B * HW * C, 256), 256, 0, (cudaStream_t)stream>>>(act, B, HW, C, flat);
  EAVIT_LAUNCH_OK();
  return EAVIT_OK;
}

int eavit_col2im_lrelu(const void* dcol, const void* act, int B, int H, int W, int C, int KH, int KW, int stride, void* din,
                       void* stream) {
  EAVIT_CHECK_ARG(dcol && act && din && B > 0 && H >= KH && W >= KW && stride > 0);
  const int OH = (H - KH) / stride + 1, OW = (W - KW) / stride + 1;
  const long long total = (long long)B * H * W * C;
  col2im_lrelu_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)dcol, (const __nv_bfloat16*)act, B, H, W, C, KH, KW, stride, OH, OW, (__nv_bfloat16*)din);
  EAVIT_LAUNCH_OK();
  return EAVIT_OK;
}

int eavit_nhwc_to_flat(const void* act, int B, int HW, int C, void* flat, void* stream) {
  EAVIT_CHECK_ARG(act && flat && B > 0 && HW > 0 && C > 0);
  nhwc_to_flat_kernel<<<cdiv((long long)B * HW * C, 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)act, B, HW, C, (__nv_bfloat16*)flat);
  EAVIT_LAUNCH_OK();
  return EAVIT_OK;
}

int eavit_flat_to_nhwc_lrelu(const void* dflat, const void* act, int B, int HW, int C, void* dact, void* stream) {
  EAVIT_CHECK_ARG(dflat && act && dact && B > 0 && HW > 0 && C > 0);
  flat_to_nhwc_lrelu_kernel<<<cdiv((long long)B * HW * C, 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)dflat, (const __nv_bfloat16*)act, B, HW, C, (__nv_bfloat16*)dact);
  EAVIT_LAUNCH_OK();
  return EAVIT_OK;
}

}  // extern "C"
