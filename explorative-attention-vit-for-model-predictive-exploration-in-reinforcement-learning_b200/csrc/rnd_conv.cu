// RND target / predictor convolutions as implicit GEMM on the tcgen05 GEMM: these kernels are the
// (de)materialisation around it.  Activations are NHWC bf16 ([B*OH*OW, C] == the GEMM's [M, N] output),
// the im2col K order is torch's weight order (c, kh, kw) so Conv2d weights are used unpermuted.
//
// Reference: model.py:368-416 (conv 8/4, 4/2, 3/1 + LeakyReLU, Flatten, Linear), agents.py:333.
#include "common.cuh"

namespace eavit {

// col[(b,oy,ox), c*KH*KW + i*KW + j] = in[b, oy*s+i, ox*s+j, c]
// One CTA per (sample, output row): the KH input rows it needs are staged in shared memory with coalesced loads, the
// (c, ki, kj) -> patch offset table is computed once per CTA (no per-element integer division), and every thread
// emits 16-byte stores of 8 consecutive K entries (hi / hi / lo segments of the bf16x3 operand when split3).
template <typename InT>
__global__ void __launch_bounds__(256) im2col_kernel(const InT* __restrict__ in, const long long* __restrict__ sample_idx, int B, int H, int W, int C, int KH, int KW,
                                                     int stride, int OH, int OW, __nv_bfloat16* __restrict__ col, int split3) {
  extern __shared__ float im_sm[];
  const int K = C * KH * KW;
  const int npatch = KH * W * C;
  float* patch = im_sm;                                   // [KH][W][C]
  int* koff = reinterpret_cast<int*>(im_sm + npatch);     // [K]
  const int b = blockIdx.x / OH, oy = blockIdx.x - b * OH;
  const long long sb = sample_idx ? sample_idx[b] : (long long)b;
  const InT* src = in + ((size_t)sb * H + (size_t)oy * stride) * W * C;
  for (int i = threadIdx.x; i < npatch; i += blockDim.x) {
    if constexpr (sizeof(InT) == 4) patch[i] = __ldg(src + i); else patch[i] = __bfloat162float(src[i]);
  }
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const int c = k / (KH * KW), r = k - c * (KH * KW), ki = r / KW, kj = r - ki * KW;
    koff[k] = (ki * W + kj) * C + c;
  }
  __syncthreads();
  const int K8 = K >> 3;
  const size_t ldc = split3 ? (size_t)3 * K : (size_t)K;
  for (int e = threadIdx.x; e < OW * K8; e += blockDim.x) {
    const int ox = e / K8, k0 = (e - ox * K8) << 3;
    const float* p = patch + ox * stride * C;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = p[koff[k0 + j]];
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      hi[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
      lo[j] = pack_bf16x2(v[2 * j] - __uint_as_float(hi[j] << 16), v[2 * j + 1] - __uint_as_float(hi[j] & 0xffff0000u));
    }
    __nv_bfloat16* row = col + ((size_t)blockIdx.x * OW + ox) * ldc + k0;
    *reinterpret_cast<uint4*>(row) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    if (split3) {
      // bf16x3 operand: [hi | hi | lo] so that A3 . B3^T (B3 = [hi | lo | hi]) = hi*hi + hi*lo + lo*hi
      *reinterpret_cast<uint4*>(row + K) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(row + 2 * K) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
}

// First convolution of the CNN actor-critic backbone (model.py:110-117, commented out upstream; BASELINE configs[1]):
// the input is the NCHW frame stack [N, C, H, W], uint8 raw frames (divided by 255 here, bit-identical to
// np.float32(x) / 255.) or float32.  Same output as im2col_kernel: col[(b,oy,ox), c*KH*KW + i*KW + j].
template <typename InT>
__global__ void __launch_bounds__(256) im2col_nchw_kernel(const InT* __restrict__ in, const long long* __restrict__ sample_idx, int B, int H, int W, int C,
                                                          int KH, int KW, int stride, int OH, int OW, __nv_bfloat16* __restrict__ col, int split3) {
  extern __shared__ float im_sm[];
  const int K = C * KH * KW;
  const int npatch = C * KH * W;
  float* patch = im_sm;                                   // [C][KH][W]
  int* koff = reinterpret_cast<int*>(im_sm + npatch);     // [K]
  const int b = blockIdx.x / OH, oy = blockIdx.x - b * OH;
  const long long sb = sample_idx ? sample_idx[b] : (long long)b;
  const InT* src = in + (size_t)sb * C * H * W + (size_t)oy * stride * W;
  for (int i = threadIdx.x; i < npatch; i += blockDim.x) {
    const int c = i / (KH * W), r = i - c * (KH * W);     // r = ki * W + x: KH consecutive image rows of channel c are contiguous
    const InT v = src[(size_t)c * H * W + r];
    if constexpr (sizeof(InT) == 1) patch[i] = __fdiv_rn((float)v, 255.0f); else patch[i] = v;
  }
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const int c = k / (KH * KW), r = k - c * (KH * KW), ki = r / KW, kj = r - ki * KW;
    koff[k] = (c * KH + ki) * W + kj;
  }
  __syncthreads();
  const int K8 = K >> 3;
  const size_t ldc = split3 ? (size_t)3 * K : (size_t)K;
  for (int e = threadIdx.x; e < OW * K8; e += blockDim.x) {
    const int ox = e / K8, k0 = (e - ox * K8) << 3;
    const float* p = patch + ox * stride;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = p[koff[k0 + j]];
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      hi[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
      lo[j] = pack_bf16x2(v[2 * j] - __uint_as_float(hi[j] << 16), v[2 * j + 1] - __uint_as_float(hi[j] & 0xffff0000u));
    }
    __nv_bfloat16* row = col + ((size_t)blockIdx.x * OW + ox) * ldc + k0;
    *reinterpret_cast<uint4*>(row) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    if (split3) {
      *reinterpret_cast<uint4*>(row + K) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(row + 2 * K) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
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

// x <- act(x + bias) in place; optional bf16 copy and [hi | hi | lo] rows (see eavit_bias_act_split3)
__global__ void __launch_bounds__(256) bias_act_split3_kernel(float* __restrict__ x, long long ldx, const float* __restrict__ bias,
                                                              int act, __nv_bfloat16* __restrict__ out16,
                                                              __nv_bfloat16* __restrict__ out3, int R, int K) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)R * (K / 2)) return;
  const int k2 = (int)(i % (K / 2));
  const long long r = i / (K / 2);
  float2 v = *reinterpret_cast<const float2*>(x + r * ldx + 2 * k2);
  if (bias != nullptr) { v.x += bias[2 * k2]; v.y += bias[2 * k2 + 1]; }
  if (act == EAVIT_ACT_RELU) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); }
  else if (act == EAVIT_ACT_LRELU) { v.x = v.x > 0.f ? v.x : 0.01f * v.x; v.y = v.y > 0.f ? v.y : 0.01f * v.y; }
  *reinterpret_cast<float2*>(x + r * ldx + 2 * k2) = v;
  const uint32_t hi = pack_bf16x2(v.x, v.y);
  if (out16 != nullptr) *reinterpret_cast<uint32_t*>(out16 + (size_t)r * K + 2 * k2) = hi;
  if (out3 != nullptr) {
    const float2 h = unpack_bf16x2(hi);
    const uint32_t lo = pack_bf16x2(v.x - h.x, v.y - h.y);
    __nv_bfloat16* row = out3 + (size_t)r * 3 * K;
    *reinterpret_cast<uint32_t*>(row + 2 * k2) = hi;
    *reinterpret_cast<uint32_t*>(row + K + 2 * k2) = hi;
    *reinterpret_cast<uint32_t*>(row + 2 * K + 2 * k2) = lo;
  }
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
// One CTA per (sample, input row): the <= ceil(KH/s) output rows of dcol that touch it are staged in shared memory with
// coalesced 16-byte loads (the direct gather read 2 bytes per 32-byte sector), then every thread sums its taps.
__global__ void __launch_bounds__(256) col2im_lrelu_kernel(const __nv_bfloat16* __restrict__ dcol, const __nv_bfloat16* __restrict__ act,
                                                           int B, int H, int W, int C, int KH, int KW, int stride, int OH,
                                                           int OW, __nv_bfloat16* __restrict__ din, float slope) {
  extern __shared__ __align__(16) unsigned char c2_sm_raw[];
  __nv_bfloat16* rows = reinterpret_cast<__nv_bfloat16*>(c2_sm_raw);     // [n_oy][OW][K]
  const int K = C * KH * KW;
  const int b = blockIdx.x / H, y = blockIdx.x - b * H;
  int oy_lo = y - KH + 1;
  oy_lo = oy_lo <= 0 ? 0 : (oy_lo + stride - 1) / stride;
  int oy_hi = y / stride;
  if (oy_hi > OH - 1) oy_hi = OH - 1;
  const int n_oy = oy_hi - oy_lo + 1;                                      // may be <= 0 at the bottom border
  if (n_oy > 0) {
    const uint4* src = reinterpret_cast<const uint4*>(dcol + ((size_t)b * OH + oy_lo) * OW * K);
    uint4* dst = reinterpret_cast<uint4*>(rows);
    const int n16 = n_oy * OW * K / 8;
    for (int i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = __ldg(src + i);
  }
  __syncthreads();
  const size_t base = ((size_t)b * H + y) * W * C;
  for (int e = threadIdx.x; e < W * C; e += blockDim.x) {
    const int x = e / C, c = e - x * C;
    int ox_lo = x - KW + 1;
    ox_lo = ox_lo <= 0 ? 0 : (ox_lo + stride - 1) / stride;
    int ox_hi = x / stride;
    if (ox_hi > OW - 1) ox_hi = OW - 1;
    float acc = 0.f;
    for (int oy = oy_lo; oy <= oy_hi; ++oy) {
      const int ki = y - oy * stride;
      const __nv_bfloat16* r = rows + (size_t)(oy - oy_lo) * OW * K + c * KH * KW + ki * KW;
      for (int ox = ox_lo; ox <= ox_hi; ++ox) acc += __bfloat162float(r[ox * K + (x - ox * stride)]);
    }
    const float a = __bfloat162float(act[base + e]);
    din[base + e] = __float2bfloat16(a > 0.f ? acc : slope * acc);
  }
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
                                                                 int B, int HW, int C, __nv_bfloat16* __restrict__ dact, float slope) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * HW * C) return;
  const int c = (int)(i % C), p = (int)((i / C) % HW);
  const long long b = i / ((long long)HW * C);
  const float g = __bfloat162float(dflat[((size_t)b * C + c) * HW + p]);
  const float a = __bfloat162float(act[i]);
  dact[i] = __float2bfloat16(a > 0.f ? g : slope * g);
}

// out = bf16(dy * act'(y)) for an activation applied OUTSIDE a GEMM epilogue (the ReLU that ends the CNN backbone)
__global__ void __launch_bounds__(256) act_bwd_bf16_kernel(const float* __restrict__ dy, const float* __restrict__ y, float slope,
                                                           __nv_bfloat16* __restrict__ out, long long n2) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n2) return;
  const float2 g = reinterpret_cast<const float2*>(dy)[i], a = reinterpret_cast<const float2*>(y)[i];
  reinterpret_cast<uint32_t*>(out)[i] = pack_bf16x2(a.x > 0.f ? g.x : slope * g.x, a.y > 0.f ? g.y : slope * g.y);
}

}  // namespace eavit

using namespace eavit;

extern "C" {

int eavit_im2col(const void* in, int in_dtype, const long long* sample_idx, int B, int H, int W, int C, int KH, int KW, int stride, void* col,
                 int split3, void* stream) {
  EAVIT_CHECK_ARG(in && col && B > 0 && H >= KH && W >= KW && stride > 0 && (C * KH * KW) % 8 == 0);
  EAVIT_CHECK_ARG((reinterpret_cast<uintptr_t>(col) & 15) == 0);
  const int OH = (H - KH) / stride + 1, OW = (W - KW) / stride + 1;
  const size_t smem = ((size_t)KH * W * C + (size_t)C * KH * KW) * 4;
  EAVIT_CHECK_ARG(smem <= 48 * 1024);
  const int grid = B * OH;
  cudaStream_t st = (cudaStream_t)stream;
  if (in_dtype == EAVIT_F32) im2col_kernel<float><<<grid, 256, smem, st>>>((const float*)in, sample_idx, B, H, W, C, KH, KW, stride, OH, OW, (__nv_bfloat16*)col, split3);
  else if (in_dtype == EAVIT_BF16) im2col_kernel<__nv_bfloat16><<<grid, 256, smem, st>>>((const __nv_bfloat16*)in, sample_idx, B, H, W, C, KH, KW, stride, OH, OW, (__nv_bfloat16*)col, split3);
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

int eavit_bias_act_split3(float* x, long long ldx, const float* bias, int act, void* out_bf16, void* out3_bf16, int R, int K,
                          void* stream) {
  EAVIT_CHECK_ARG(x && R > 0 && K > 0 && K % 2 == 0 && ldx % 2 == 0);
  EAVIT_CHECK_ARG(act == EAVIT_ACT_NONE || act == EAVIT_ACT_RELU || act == EAVIT_ACT_LRELU);
  bias_act_split3_kernel<<<cdiv((long long)R * (K / 2), 256), 256, 0, (cudaStream_t)stream>>>(
      x, ldx, bias, act, (__nv_bfloat16*)out_bf16, (__nv_bfloat16*)out3_bf16, R, K);
  EAVIT_LAUNCH_OK();
  return EAVIT_OK;
}

int eavit_nhwc_to_flat_f32(const float* act, int B, int HW, int C, float* flat, void* stream) {
  EAVIT_CHECK_ARG(act && flat && B > 0 && HW > 0 && C > 0);
  nhwc_to_flat_f32_kernel<<<cdiv((long long)B * HW * C, 256), 256, 0, (cudaStream_t)stream>>>(act, B, HW, C, flat);
  EAVIT_LAUNCH_OK();
  return EAVIT_OK;
}

int eavit_im2col_nchw(const void* in, int in_dtype, const long long* sample_idx, int B, int H, int W, int C, int KH, int KW, int stride,
                      void* col, int split3, void* stream) {
  EAVIT_CHECK_ARG(in && col && B > 0 && H >= KH && W >= KW && stride > 0 && C > 0 && (C * KH * KW) % 8 == 0);
  EAVIT_CHECK_ARG((reinterpret_cast<uintptr_t>(col) & 15) == 0);
  const int OH = (H - KH) / stride + 1, OW = (W - KW) / stride + 1;
  const size_t smem = ((size_t)KH * W * C + (size_t)C * KH * KW) * 4;
  EAVIT_CHECK_ARG(smem <= 48 * 1024);
  const int grid = B * OH;
  cudaStream_t st = (cudaStream_t)stream;
  if (in_dtype == EAVIT_F32) im2col_nchw_kernel<float><<<grid, 256, smem, st>>>((const float*)in, sample_idx, B, H, W, C, KH, KW, stride, OH, OW, (__nv_bfloat16*)col, split3);
  else if (in_dtype == EAVIT_U8) im2col_nchw_kernel<uint8_t><<<grid, 256, smem, st>>>((const uint8_t*)in, sample_idx, B, H, W, C, KH, KW, stride, OH, OW, (__nv_bfloat16*)col, split3);
  else { set_error("im2col_nchw: bad dtype %d", in_dtype); return EAVIT_EINVAL; }
  EAVIT_LAUNCH_OK();
  return EAVIT_OK;
}

static int col2im_launch(const void* dcol, const void* act, int B, int H, int W, int C, int KH, int KW, int stride, void* din, float slope,
                         void* stream);
int eavit_col2im_lrelu(const void* dcol, const void* act, int B, int H, int W, int C, int KH, int KW, int stride, void* din,
                       void* stream) {
  return col2im_launch(dcol, act, B, H, W, C, KH, KW, stride, din, 0.01f, stream);
}
int eavit_col2im_act(const void* dcol, const void* act, int B, int H, int W, int C, int KH, int KW, int stride, void* din, float slope,
                     void* stream) {
  return col2im_launch(dcol, act, B, H, W, C, KH, KW, stride, din, slope, stream);
}
static int col2im_launch(const void* dcol, const void* act, int B, int H, int W, int C, int KH, int KW, int stride, void* din, float slope,
                         void* stream) {
  EAVIT_CHECK_ARG(dcol && act && din && B > 0 && H >= KH && W >= KW && stride > 0);
  const int OH = (H - KH) / stride + 1, OW = (W - KW) / stride + 1;
  const int K = C * KH * KW;
  EAVIT_CHECK_ARG(K % 8 == 0 && (reinterpret_cast<uintptr_t>(dcol) & 15) == 0);
  const size_t smem = (size_t)((KH + stride - 1) / stride) * OW * K * 2;
  EAVIT_CHECK_ARG(smem <= 48 * 1024);
  col2im_lrelu_kernel<<<B * H, 256, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)dcol, (const __nv_bfloat16*)act, B, H, W, C, KH, KW, stride, OH, OW, (__nv_bfloat16*)din, slope);
  EAVIT_LAUNCH_OK();
  return EAVIT_OK;
}

int eavit_nhwc_to_flat(const void* act, int B, int HW, int C, void* flat, void* stream) {
  EAVIT_CHECK_ARG(act && flat && B > 0 && HW > 0 && C > 0);
  nhwc_to_flat_kernel<<<cdiv((long long)B * HW * C, 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)act, B, HW, C, (__nv_bfloat16*)flat);
  EAVIT_LAUNCH_OK();
  return EAVIT_OK;
}

int eavit_flat_to_nhwc_act(const void* dflat, const void* act, int B, int HW, int C, void* dact, float slope, void* stream) {
  EAVIT_CHECK_ARG(dflat && act && dact && B > 0 && HW > 0 && C > 0);
  flat_to_nhwc_lrelu_kernel<<<cdiv((long long)B * HW * C, 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)dflat, (const __nv_bfloat16*)act, B, HW, C, (__nv_bfloat16*)dact, slope);
  EAVIT_LAUNCH_OK();
  return EAVIT_OK;
}
int eavit_flat_to_nhwc_lrelu(const void* dflat, const void* act, int B, int HW, int C, void* dact, void* stream) {
  return eavit_flat_to_nhwc_act(dflat, act, B, HW, C, dact, 0.01f, stream);
}

int eavit_act_bwd_bf16(const float* dy, const float* y, float slope, void* out_bf16, long long n, void* stream) {
  EAVIT_CHECK_ARG(dy && y && out_bf16 && n > 0 && n % 2 == 0);
  act_bwd_bf16_kernel<<<cdiv(n / 2, 256), 256, 0, (cudaStream_t)stream>>>(dy, y, slope, (__nv_bfloat16*)out_bf16, n / 2);
  EAVIT_LAUNCH_OK();
  return EAVIT_OK;
}

}  // extern "C"
