// Multi-head self-attention over short sequences (S <= 256: 196/197 patch-6 tokens, 50 patch-12 tokens).
// One CTA per (sequence, head): the whole head's Q/K/V tile lives in shared memory, scores never touch HBM
// (the reference materialises [B,8,S,S] fp32 scores -- 636 MB per layer at B=512).
//
// Layout: qkv bf16 [T, 3*H*Dh] = per token  q(h d) | k(h d) | v(h d)   (vit.py:63-64 chunk + rearrange,
// HF view(b, n, heads, d)); out bf16 [T, H*Dh]  ('b h n d -> b n (h d)', vit.py:72); lse fp32 [T, H].
// Sequences are described by seq_start[nseq+1] (token offsets), so the 196- and 197-token passes of the
// explorative / exploitative sequences run in one launch.
//
// v1 arithmetic is fp32 on CUDA cores from bf16 operands (softmax in fp32, exact expf).
#include "common.cuh"

namespace eavit {

constexpr int ATT_WARPS = 8;
constexpr int ATT_MAXS = 256;

template <int DH>
__global__ void __launch_bounds__(ATT_WARPS * 32) attention_fwd_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                                     const int* __restrict__ seq_start, int H,
                                                                     float scale, __nv_bfloat16* __restrict__ out,
                                                                     float* __restrict__ lse, int SM) {
  constexpr int LD = DH + 1;
  constexpr int R = DH / 32;
  extern __shared__ float sm[];
  const int seq = blockIdx.x / H, h = blockIdx.x % H;
  const int t0 = seq_start[seq], S = seq_start[seq + 1] - t0;
  float* sK = sm;                         // [S][LD]
  float* sV = sK + SM * LD;               // [S][LD]
  float* sP = sV + SM * LD;               // [ATT_WARPS][SM]
  float* sQ = sP + ATT_WARPS * SM;        // [ATT_WARPS][DH]
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ldq = 3 * H * DH;
  // stage K, V (bf16 -> fp32), two elements per thread-iteration
  for (int i = threadIdx.x; i < S * (DH / 2); i += blockDim.x) {
    const int j = i / (DH / 2), d2 = i % (DH / 2);
    const __nv_bfloat16* base = qkv + (size_t)(t0 + j) * ldq + h * DH + 2 * d2;
    const float2 k = unpack_bf16x2(__ldg(reinterpret_cast<const uint32_t*>(base + H * DH)));
    const float2 v = unpack_bf16x2(__ldg(reinterpret_cast<const uint32_t*>(base + 2 * H * DH)));
    sK[j * LD + 2 * d2] = k.x; sK[j * LD + 2 * d2 + 1] = k.y;
    sV[j * LD + 2 * d2] = v.x; sV[j * LD + 2 * d2 + 1] = v.y;
  }
  __syncthreads();
  float* myP = sP + w * SM;
  float* myQ = sQ + w * DH;
  for (int i = w; i < S; i += ATT_WARPS) {
    const __nv_bfloat16* qrow = qkv + (size_t)(t0 + i) * ldq + h * DH;
#pragma unroll
    for (int r = 0; r < R; ++r) myQ[lane + 32 * r] = __bfloat162float(qrow[lane + 32 * r]) * scale;
    __syncwarp();
    float sc[ATT_MAXS / 32];
    float mx = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < ATT_MAXS / 32; ++jj) {
      const int j = lane + 32 * jj;
      float a = -INFINITY;
      if (j < S) {
        a = 0.f;
        const float* kr = sK + j * LD;
#pragma unroll
        for (int d = 0; d < DH; ++d) a += myQ[d] * kr[d];
      }
      sc[jj] = a;
      mx = fmaxf(mx, a);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int jj = 0; jj < ATT_MAXS / 32; ++jj) {
      const int j = lane + 32 * jj;
      const float e = (j < S) ? __expf(sc[jj] - mx) : 0.f;
      sum += e;
      if (j < S) myP[j] = e;
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    __syncwarp();
    float acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0.f;
    for (int j = 0; j < S; ++j) {
      const float pj = myP[j];
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r] += pj * sV[j * LD + lane + 32 * r];
    }
    __nv_bfloat16* orow = out + (size_t)(t0 + i) * (H * DH) + h * DH;
#pragma unroll
    for (int r = 0; r < R; ++r) orow[lane + 32 * r] = __float2bfloat16(acc[r] * inv);
    if (lane == 0 && lse != nullptr) lse[(size_t)(t0 + i) * H + h] = mx + __logf(sum);
    __syncwarp();
  }
}

// Backward: recompute P from (q, k, lse).  Phase A (per query row): dQ.  Phase B (per key row): dK, dV.
//   D_i = sum_j P_ij dP_ij (== sum_d dO_i O_i) ; dS = P * (dP - D) ; dQ = scale dS K ; dK = scale dS^T Q ; dV = P^T dO
template <int DH>
__global__ void __launch_bounds__(ATT_WARPS * 32) attention_bwd_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                                     const __nv_bfloat16* __restrict__ o,
                                                                     const __nv_bfloat16* __restrict__ dout,
                                                                     const float* __restrict__ lse,
                                                                     const int* __restrict__ seq_start, int H,
                                                                     float scale, __nv_bfloat16* __restrict__ dqkv, int SM) {
  constexpr int LD = DH + 1;
  constexpr int R = DH / 32;
  extern __shared__ float sm[];
  const int seq = blockIdx.x / H, h = blockIdx.x % H;
  const int t0 = seq_start[seq], S = seq_start[seq + 1] - t0;
  float* sQ = sm;                          // [S][LD]  (pre-scaled by `scale`)
  float* sK = sQ + SM * LD;
  float* sV = sK + SM * LD;
  float* sdO = sV + SM * LD;
  float* sL = sdO + SM * LD;               // [S] lse
  float* sD = sL + SM;                     // [S] D_i
  float* sA = sD + SM;                     // [ATT_WARPS][SM]  p or ds scratch
  float* sB = sA + ATT_WARPS * SM;         // [ATT_WARPS][SM]
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ldq = 3 * H * DH, ldo = H * DH;
  for (int i = threadIdx.x; i < S * (DH / 2); i += blockDim.x) {
    const int j = i / (DH / 2), d2 = i % (DH / 2);
    const __nv_bfloat16* base = qkv + (size_t)(t0 + j) * ldq + h * DH + 2 * d2;
    const float2 q = unpack_bf16x2(__ldg(reinterpret_cast<const uint32_t*>(base)));
    const float2 k = unpack_bf16x2(__ldg(reinterpret_cast<const uint32_t*>(base + H * DH)));
    const float2 v = unpack_bf16x2(__ldg(reinterpret_cast<const uint32_t*>(base + 2 * H * DH)));
    const float2 g = unpack_bf16x2(__ldg(reinterpret_cast<const uint32_t*>(dout + (size_t)(t0 + j) * ldo + h * DH + 2 * d2)));
    sQ[j * LD + 2 * d2] = q.x * scale; sQ[j * LD + 2 * d2 + 1] = q.y * scale;
    sK[j * LD + 2 * d2] = k.x; sK[j * LD + 2 * d2 + 1] = k.y;
    sV[j * LD + 2 * d2] = v.x; sV[j * LD + 2 * d2 + 1] = v.y;
    sdO[j * LD + 2 * d2] = g.x; sdO[j * LD + 2 * d2 + 1] = g.y;
  }
  for (int i = threadIdx.x; i < S; i += blockDim.x) sL[i] = lse[(size_t)(t0 + i) * H + h];
  __syncthreads();
  float* myA = sA + w * SM;
  float* myB = sB + w * SM;
  // ---- phase A: dQ_i = scale * sum_j dS_ij K_j.  D_i = sum_j P_ij dP_ij is taken from the SAME recomputed P
  // that multiplies (dP - D), so the softmax Jacobian is applied consistently (no bf16-rounded O involved).
  for (int i = w; i < S; i += ATT_WARPS) {
    const float* qi = sQ + i * LD;
    const float* gi = sdO + i * LD;
    const float li = sL[i];
    float dsum = 0.f;
#pragma unroll
    for (int jj = 0; jj < ATT_MAXS / 32; ++jj) {
      const int j = lane + 32 * jj;
      if (j < S) {
        float s = 0.f, dp = 0.f;
        const float* kr = sK + j * LD;
        const float* vr = sV + j * LD;
#pragma unroll
        for (int d = 0; d < DH; ++d) { s += qi[d] * kr[d]; dp += gi[d] * vr[d]; }
        const float p = __expf(s - li);
        myA[j] = p;
        myB[j] = dp;
        dsum += p * dp;
      }
    }
    const float Di = warp_sum(dsum);
    if (lane == 0) sD[i] = Di;
    __syncwarp();
    float acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0.f;
    for (int j = 0; j < S; ++j) {
      const float ds = myA[j] * (myB[j] - Di);
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r] += ds * sK[j * LD + lane + 32 * r];
    }
    __nv_bfloat16* dq = dqkv + (size_t)(t0 + i) * ldq + h * DH;
#pragma unroll
    for (int r = 0; r < R; ++r) dq[lane + 32 * r] = __float2bfloat16(acc[r] * scale);
    __syncwarp();
  }
  __syncthreads();     // every D_i is needed by phase B
  // ---- phase B: dK_j = sum_i dS_ij (scale Q_i) ; dV_j = sum_i P_ij dO_i
  for (int j = w; j < S; j += ATT_WARPS) {
    const float* kj = sK + j * LD;
    const float* vj = sV + j * LD;
#pragma unroll
    for (int ii = 0; ii < ATT_MAXS / 32; ++ii) {
      const int i = lane + 32 * ii;
      if (i < S) {
        float s = 0.f, dp = 0.f;
        const float* qr = sQ + i * LD;
        const float* gr = sdO + i * LD;
#pragma unroll
        for (int d = 0; d < DH; ++d) { s += qr[d] * kj[d]; dp += gr[d] * vj[d]; }
        const float p = __expf(s - sL[i]);
        myA[i] = p;
        myB[i] = p * (dp - sD[i]);
      }
    }
    __syncwarp();
    float ak[R], av[R];
#pragma unroll
    for (int r = 0; r < R; ++r) { ak[r] = 0.f; av[r] = 0.f; }
    for (int i = 0; i < S; ++i) {
      const float p = myA[i], ds = myB[i];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        ak[r] += ds * sQ[i * LD + lane + 32 * r];     // sQ already carries `scale`
        av[r] += p * sdO[i * LD + lane + 32 * r];
      }
    }
    __nv_bfloat16* dk = dqkv + (size_t)(t0 + j) * ldq + H * DH + h * DH;
    __nv_bfloat16* dv = dk + H * DH;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      dk[lane + 32 * r] = __float2bfloat16(ak[r]);
      dv[lane + 32 * r] = __float2bfloat16(av[r]);
    }
    __syncwarp();
  }
}

static size_t att_fwd_smem(int DH, int SM) { return sizeof(float) * ((size_t)2 * SM * (DH + 1) + ATT_WARPS * SM + ATT_WARPS * DH); }
static size_t att_bwd_smem(int DH, int SM) { return sizeof(float) * ((size_t)4 * SM * (DH + 1) + 2 * SM + 2 * ATT_WARPS * SM); }

}  // namespace eavit

using namespace eavit;

extern "C" {

int eavit_attention_fwd(const void* qkv, const int* seq_start, int nseq, int max_len, int H, int Dh, float scale,
                        void* out, float* lse, void* stream) {
  EAVIT_CHECK_ARG(qkv && seq_start && out && nseq > 0 && H > 0);
  EAVIT_CHECK_ARG(max_len > 0 && max_len <= ATT_MAXS);
  EAVIT_CHECK_ARG(Dh == 32 || Dh == 64);
  cudaStream_t st = (cudaStream_t)stream;
  const int SM = ((max_len + 31) / 32) * 32;
  const size_t smem = att_fwd_smem(Dh, SM);
  EAVIT_CHECK_ARG(smem <= 227 * 1024);
  if (Dh == 32) {
    static bool done = false;
    if (!done) { EAVIT_CUDA(cudaFuncSetAttribute(attention_fwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); done = true; }
    attention_fwd_kernel<32><<<nseq * H, ATT_WARPS * 32, smem, st>>>((const __nv_bfloat16*)qkv, seq_start, H, scale, (__nv_bfloat16*)out, lse, SM);
  } else {
    static bool done = false;
    if (!done) { EAVIT_CUDA(cudaFuncSetAttribute(attention_fwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); done = true; }
    attention_fwd_kernel<64><<<nseq * H, ATT_WARPS * 32, smem, st>>>((const __nv_bfloat16*)qkv, seq_start, H, scale, (__nv_bfloat16*)out, lse, SM);
  }
  EAVIT_LAUNCH_OK();
  return EAVIT_OK;
}

int eavit_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse, const int* seq_start,
                        int nseq, int max_len, int H, int Dh, float scale, void* dqkv, void* stream) {
  EAVIT_CHECK_ARG(qkv && out && dout && lse && seq_start && dqkv && nseq > 0 && H > 0);
  EAVIT_CHECK_ARG(max_len > 0 && max_len <= ATT_MAXS);
  EAVIT_CHECK_ARG(Dh == 32 || Dh == 64);
  cudaStream_t st = (cudaStream_t)stream;
  const int SM = ((max_len + 31) / 32) * 32;
  const size_t smem = att_bwd_smem(Dh, SM);
  EAVIT_CHECK_ARG(smem <= 227 * 1024);
  if (Dh == 32) {
    static bool done = false;
    if (!done) { EAVIT_CUDA(cudaFuncSetAttribute(attention_bwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); done = true; }
    attention_bwd_kernel<32><<<nseq * H, ATT_WARPS * 32, smem, st>>>((const __nv_bfloat16*)qkv, (const __nv_bfloat16*)out, (const __nv_bfloat16*)dout, lse, seq_start, H, scale, (__nv_bfloat16*)dqkv, SM);
  } else {
    static bool done = false;
    if (!done) { EAVIT_CUDA(cudaFuncSetAttribute(attention_bwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); done = true; }
    attention_bwd_kernel<64><<<nseq * H, ATT_WARPS * 32, smem, st>>>((const __nv_bfloat16*)qkv, (const __nv_bfloat16*)out, (const __nv_bfloat16*)dout, lse, seq_start, H, scale, (__nv_bfloat16*)dqkv, SM);
  }
  EAVIT_LAUNCH_OK();
  return EAVIT_OK;
}

}  // extern "C"
