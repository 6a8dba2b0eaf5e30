// Attention of the pooled query only (last transformer layer).
//
// The reference pools token 0 of each sequence after the last layer (x[:, 0], vit.py:162; out[0][:, 0, :], model.py:316)
// while computing every token of that layer.  Everything the last layer does for query rows >= 1 -- their attention
// output, out-projection, LayerNorm and MLP -- never reaches the loss, in either direction, so engine.ViTEncoder runs
// the last layer's token-wise part on the pooled rows only and its attention with ONE query row per (sequence, head):
//
//   forward :  s_j = q_0 . k_j * scale,  p = softmax(s),  o_0 = sum_j (p_j m_j) v_j                 (m = dropout mask)
//   backward:  dp_j = do_0 . v_j,  D = sum_j p_j m_j dp_j,  ds_j = p_j (m_j dp_j - D)
//              dq_0 = scale sum_j ds_j k_j,   dk_j = scale ds_j q_0,   dv_j = p_j m_j do_0,   dq_{i>0} = 0
//
// K and V still come from every token, so the gradient flowing back into the residual stream is dense and layers below
// are untouched.  This is a matrix-vector problem (2 * S * Dh MACs per head), so it runs on CUDA cores: one warp per
// (sequence, head), lanes over keys for the scores, lanes over head dimensions for the weighted sums; the kernels are
// bound by streaming K / V (forward: 206 MB at the cfg3 shape) and writing dqkv (backward: 309 MB).
#include "common.cuh"

namespace eavit {

constexpr int R0_WARPS = 8;
constexpr int R0_MAXPL = 16;          // keys per lane: sequences up to 512 tokens

// sum_d row[d] * q[d]; row = DH bf16 in global memory (one key / value row), q = DH floats in shared memory (all lanes of a
// warp read the same q element in the same iteration: broadcast, conflict-free)
template <int DH>
__device__ __forceinline__ float dot_row(const __nv_bfloat16* row, const float* q) {
  float a0 = 0.f, a1 = 0.f;
#pragma unroll
  for (int c = 0; c < DH / 8; ++c) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(row) + c);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      a0 = fmaf(__uint_as_float(w[i] << 16), q[c * 8 + 2 * i], a0);
      a1 = fmaf(__uint_as_float(w[i] & 0xffff0000u), q[c * 8 + 2 * i + 1], a1);
    }
  }
  return a0 + a1;
}

// normalised p_j = softmax_j(q_0 . k_j * scale) for this lane's keys j = lane + 32 i
template <int DH>
__device__ __forceinline__ void row0_softmax(const __nv_bfloat16* kbase, int ldq, int S, const float* q, float scale, int lane,
                                             float* p) {
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < R0_MAXPL; ++i) {
    const int j = lane + 32 * i;
    p[i] = -INFINITY;
    if (j < S) { p[i] = dot_row<DH>(kbase + (size_t)j * ldq, q) * scale; mx = fmaxf(mx, p[i]); }
  }
  mx = warp_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < R0_MAXPL; ++i) {
    const int j = lane + 32 * i;
    p[i] = (j < S) ? __expf(p[i] - mx) : 0.f;
    sum += p[i];
  }
  const float inv = 1.0f / warp_sum(sum);
#pragma unroll
  for (int i = 0; i < R0_MAXPL; ++i) p[i] *= inv;
}

// dropout factor (0 or 1/(1-p)) of attention-probability element (token row key `rk`, column h*256 + j) -- the same
// element the tensor-core kernels use (vit.py:70)
__device__ __forceinline__ float row0_mask(const DropCfg& drop, uint32_t rk, int h, int j) {
  if (drop.thresh == 0) return 1.f;
  const uint32_t c = (uint32_t)h * 256u + (uint32_t)j;
  const uint32_t bits = drop_bits(rk, c >> 1);
  return (c & 1) ? drop_odd(drop, bits) : drop_even(drop, bits);
}

template <int DH>
__global__ void __launch_bounds__(R0_WARPS * 32) attention_row0_fwd_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                                          const int* __restrict__ seq_start, int nseq, int H,
                                                                          float scale, __nv_bfloat16* __restrict__ out0,
                                                                          const DropCfg drop) {
  constexpr int DPL = DH / 32;                     // head dimensions per lane
  __shared__ float s_q[R0_WARPS][DH];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * R0_WARPS + w;
  if (item >= nseq * H) return;
  const int seq = item / H, h = item - seq * H;
  const int t0 = seq_start[seq], S = seq_start[seq + 1] - t0;
  const int ldq = 3 * H * DH;
  const __nv_bfloat16* qrow = qkv + (size_t)t0 * ldq + h * DH;
#pragma unroll
  for (int e = 0; e < DPL; ++e) s_q[w][lane + 32 * e] = __bfloat162float(qrow[lane + 32 * e]);
  __syncwarp();
  float p[R0_MAXPL];
  row0_softmax<DH>(qrow + H * DH, ldq, S, s_q[w], scale, lane, p);
  const uint32_t rk = drop_row_key(drop, (uint32_t)t0);
#pragma unroll
  for (int i = 0; i < R0_MAXPL; ++i) p[i] *= row0_mask(drop, rk, h, lane + 32 * i);
  // o_0[d] = sum_j p_j m_j v_j[d]: lanes over d, p_j broadcast from its owner lane
  const __nv_bfloat16* vbase = qrow + 2 * H * DH;
  float o[DPL];
#pragma unroll
  for (int e = 0; e < DPL; ++e) o[e] = 0.f;
#pragma unroll
  for (int i = 0; i < R0_MAXPL; ++i) {
    if (32 * i >= S) break;
    const int n = min(32, S - 32 * i);
    for (int l = 0; l < n; ++l) {
      const float pj = __shfl_sync(0xffffffffu, p[i], l);
      const __nv_bfloat16* vr = vbase + (size_t)(32 * i + l) * ldq;
#pragma unroll
      for (int e = 0; e < DPL; ++e) o[e] = fmaf(pj, __bfloat162float(vr[lane + 32 * e]), o[e]);
    }
  }
#pragma unroll
  for (int e = 0; e < DPL; ++e) out0[(size_t)seq * H * DH + h * DH + lane + 32 * e] = __float2bfloat16(o[e]);
}

template <int DH>
__global__ void __launch_bounds__(R0_WARPS * 32) attention_row0_bwd_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                                          const __nv_bfloat16* __restrict__ dout0,
                                                                          const int* __restrict__ seq_start, int nseq, int H,
                                                                          float scale, __nv_bfloat16* __restrict__ dqkv,
                                                                          const DropCfg drop) {
  constexpr int DPL = DH / 32;
  __shared__ float s_q[R0_WARPS][DH], s_g[R0_WARPS][DH];             // q_0 and do_0 of this warp's (sequence, head)
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * R0_WARPS + w;
  if (item >= nseq * H) return;
  const int seq = item / H, h = item - seq * H;
  const int t0 = seq_start[seq], S = seq_start[seq + 1] - t0;
  const int ldq = 3 * H * DH;
  const __nv_bfloat16* qrow = qkv + (size_t)t0 * ldq + h * DH;
  const __nv_bfloat16* kbase = qrow + H * DH;
  const __nv_bfloat16* vbase = qrow + 2 * H * DH;
  float ql[DPL], gl[DPL];                                            // this lane's dimensions of q_0 / do_0
#pragma unroll
  for (int e = 0; e < DPL; ++e) {
    ql[e] = __bfloat162float(qrow[lane + 32 * e]);
    gl[e] = __bfloat162float(dout0[(size_t)seq * H * DH + h * DH + lane + 32 * e]);
    s_q[w][lane + 32 * e] = ql[e];
    s_g[w][lane + 32 * e] = gl[e];
  }
  __syncwarp();
  float p[R0_MAXPL], pm[R0_MAXPL], ds[R0_MAXPL];
  row0_softmax<DH>(kbase, ldq, S, s_q[w], scale, lane, p);
  const uint32_t rk = drop_row_key(drop, (uint32_t)t0);
  float dsum = 0.f;
#pragma unroll
  for (int i = 0; i < R0_MAXPL; ++i) {
    const int j = lane + 32 * i;
    const float m = row0_mask(drop, rk, h, j);
    const float dp = (j < S) ? dot_row<DH>(vbase + (size_t)j * ldq, s_g[w]) : 0.f;   // gradient w.r.t. the dropped p_j m_j
    pm[i] = p[i] * m;                                                 // weight of v_j in o_0 = weight of do_0 in dv_j
    ds[i] = m * dp;                                                   // gradient w.r.t. the softmax output p_j
    dsum = fmaf(p[i], ds[i], dsum);
  }
  const float D = warp_sum(dsum);
#pragma unroll
  for (int i = 0; i < R0_MAXPL; ++i) ds[i] = p[i] * (ds[i] - D) * scale;   // softmax Jacobian; scale folded in
  // lanes over head dimensions: dq_0 accumulation; dk_j / dv_j rows written as contiguous 2*DH-byte segments
  __nv_bfloat16* dq = dqkv + (size_t)t0 * ldq + h * DH;
  __nv_bfloat16* dk = dq + H * DH;
  __nv_bfloat16* dv = dq + 2 * H * DH;
  float dq0[DPL];
#pragma unroll
  for (int e = 0; e < DPL; ++e) dq0[e] = 0.f;
  const __nv_bfloat16 zero = __float2bfloat16(0.f);
#pragma unroll
  for (int i = 0; i < R0_MAXPL; ++i) {
    if (32 * i >= S) break;
    const int n = min(32, S - 32 * i);
    for (int l = 0; l < n; ++l) {
      const float dsj = __shfl_sync(0xffffffffu, ds[i], l);
      const float pmj = __shfl_sync(0xffffffffu, pm[i], l);
      const size_t ro = (size_t)(32 * i + l) * ldq;
#pragma unroll
      for (int e = 0; e < DPL; ++e) {
        const int d = lane + 32 * e;
        dq0[e] = fmaf(dsj, __bfloat162float(kbase[ro + d]), dq0[e]);
        dk[ro + d] = __float2bfloat16(dsj * ql[e]);
        dv[ro + d] = __float2bfloat16(pmj * gl[e]);
        if (32 * i + l > 0) dq[ro + d] = zero;                        // query rows >= 1 carry no gradient
      }
    }
  }
#pragma unroll
  for (int e = 0; e < DPL; ++e) dq[lane + 32 * e] = __float2bfloat16(dq0[e]);
}

}  // namespace eavit

using namespace eavit;

extern "C" int eavit_attention_row0_fwd(const void* qkv, const int* seq_start, int nseq, int max_len, int H, int Dh, float scale,
                                        void* out0, float drop_p, unsigned long long drop_seed, void* stream) {
  EAVIT_CHECK_ARG(qkv && seq_start && out0 && nseq > 0 && H > 0 && (Dh == 32 || Dh == 64));
  EAVIT_CHECK_ARG(max_len > 0 && max_len <= 32 * R0_MAXPL && drop_p >= 0.f && drop_p < 1.f);
  EAVIT_CHECK_ARG((reinterpret_cast<uintptr_t>(qkv) & 15) == 0);
  const DropCfg drop = make_drop(drop_p, drop_seed);
  const int grid = cdiv((long long)nseq * H, R0_WARPS);
  cudaStream_t st = (cudaStream_t)stream;
  if (Dh == 32) attention_row0_fwd_kernel<32><<<grid, R0_WARPS * 32, 0, st>>>((const __nv_bfloat16*)qkv, seq_start, nseq, H, scale, (__nv_bfloat16*)out0, drop);
  else          attention_row0_fwd_kernel<64><<<grid, R0_WARPS * 32, 0, st>>>((const __nv_bfloat16*)qkv, seq_start, nseq, H, scale, (__nv_bfloat16*)out0, drop);
  EAVIT_LAUNCH_OK();
  return EAVIT_OK;
}

extern "C" int eavit_attention_row0_bwd(const void* qkv, const void* dout0, const int* seq_start, int nseq, int max_len, int H,
                                        int Dh, float scale, void* dqkv, float drop_p, unsigned long long drop_seed,
                                        void* stream) {
  EAVIT_CHECK_ARG(qkv && dout0 && seq_start && dqkv && nseq > 0 && H > 0 && (Dh == 32 || Dh == 64));
  EAVIT_CHECK_ARG(max_len > 0 && max_len <= 32 * R0_MAXPL && drop_p >= 0.f && drop_p < 1.f);
  EAVIT_CHECK_ARG((reinterpret_cast<uintptr_t>(qkv) & 15) == 0);
  const DropCfg drop = make_drop(drop_p, drop_seed);
  const int grid = cdiv((long long)nseq * H, R0_WARPS);
  cudaStream_t st = (cudaStream_t)stream;
  if (Dh == 32) attention_row0_bwd_kernel<32><<<grid, R0_WARPS * 32, 0, st>>>((const __nv_bfloat16*)qkv, (const __nv_bfloat16*)dout0, seq_start, nseq, H, scale, (__nv_bfloat16*)dqkv, drop);
  else          attention_row0_bwd_kernel<64><<<grid, R0_WARPS * 32, 0, st>>>((const __nv_bfloat16*)qkv, (const __nv_bfloat16*)dout0, seq_start, nseq, H, scale, (__nv_bfloat16*)dqkv, drop);
  EAVIT_LAUNCH_OK();
  return EAVIT_OK;
}
