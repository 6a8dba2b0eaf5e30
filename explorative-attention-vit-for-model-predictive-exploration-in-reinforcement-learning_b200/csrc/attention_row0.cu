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
// are untouched.  This is a matrix-vector problem (2 * S * Dh MACs per head), so it runs on CUDA cores and is bound by
// streaming K / V (forward: 206 MB at the cfg3 shape) and writing dqkv (backward: 309 MB).  One CTA per sequence, all
// heads together: a warp reads one key's K (or V) section as 16 bytes per lane -- whole 512-byte row pieces, every
// sector used -- with several keys in flight per warp; the DH / 8 lanes of a head combine by shuffles, the scores of all
// heads sit in shared memory, softmax runs one warp per head, and the backward writes dq / dk / dv rows as 16-byte
// stores of the same shape.
#include "common.cuh"

namespace eavit {

constexpr int R0_WARPS = 8;
constexpr int R0_THREADS = R0_WARPS * 32;
constexpr int R0_MAXLEN = 512;
constexpr int R0_UN = 4;              // keys in flight per warp

__device__ __forceinline__ void bf16x8_to_f32(const uint4& u, float* f) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) { f[2 * i] = __uint_as_float(w[i] << 16); f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
}
__device__ __forceinline__ float dot8(const uint4& u, const float* q) {
  float f[8];
  bf16x8_to_f32(u, f);
  float a0 = 0.f, a1 = 0.f;
#pragma unroll
  for (int i = 0; i < 8; i += 2) { a0 = fmaf(f[i], q[i], a0); a1 = fmaf(f[i + 1], q[i + 1], a1); }
  return a0 + a1;
}
__device__ __forceinline__ uint4 scaled8(float s, const float* q) {
  return make_uint4(pack_bf16x2(s * q[0], s * q[1]), pack_bf16x2(s * q[2], s * q[3]), pack_bf16x2(s * q[4], s * q[5]),
                    pack_bf16x2(s * q[6], s * q[7]));
}
// sum over the LPH = DH / 8 consecutive lanes that share a head
template <int DH>
__device__ __forceinline__ float head_sum(float v) {
#pragma unroll
  for (int o = DH / 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// dropout factor (0 or 1/(1-p)) of attention-probability element (token row key `rk`, column h*256 + j) -- the same
// element the tensor-core kernels use (vit.py:70)
__device__ __forceinline__ float row0_mask(const DropCfg& drop, uint32_t rk, int h, int j) {
  if (drop.thresh == 0) return 1.f;
  const uint32_t c = (uint32_t)h * 256u + (uint32_t)j;
  const uint32_t bits = drop_bits(rk, c >> 1);
  return (c & 1) ? drop_odd(drop, bits) : drop_even(drop, bits);
}

// Column chunk c (256 bf16 = 512 bytes of a K / V / Q section) x lane -> columns [c*256 + lane*8, +8), head (c*256 + lane*8) / DH.
// NCH = ceil(H * DH / 256) chunks per section (1 for the lucidrains ViT, 4 for the HF-style one); lanes past H * DH idle.
// Shared memory: sc[H][SP] scores -> probabilities (-> ds in the backward), dp[H][SP] (backward), red[R0_WARPS][H*DH].
template <int DH, int NCH, bool BWD>
__device__ __forceinline__ void row0_scores(const __nv_bfloat16* __restrict__ kbase, const __nv_bfloat16* __restrict__ vbase,
                                            int ldq, int HD, int S, int SP, const float (*q)[8], const float (*g)[8],
                                            float scale, float* sc, float* dp, int warp, int lane) {
  constexpr int LPH = DH / 8, HPC = 256 / DH;
  for (int j0 = warp; j0 < S; j0 += R0_WARPS * R0_UN) {
    uint4 kk[R0_UN][NCH], vv[R0_UN][NCH];
#pragma unroll
    for (int u = 0; u < R0_UN; ++u) {
      const int j = j0 + u * R0_WARPS;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        kk[u][c] = make_uint4(0, 0, 0, 0);
        if (BWD) vv[u][c] = make_uint4(0, 0, 0, 0);
        if (j < S && c * 256 + lane * 8 < HD) {
          kk[u][c] = __ldg(reinterpret_cast<const uint4*>(kbase + (size_t)j * ldq + c * 256) + lane);
          if (BWD) vv[u][c] = __ldg(reinterpret_cast<const uint4*>(vbase + (size_t)j * ldq + c * 256) + lane);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < R0_UN; ++u) {
      const int j = j0 + u * R0_WARPS;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const float s = head_sum<DH>(dot8(kk[u][c], q[c]));
        float d = 0.f;
        if (BWD) d = head_sum<DH>(dot8(vv[u][c], g[c]));
        if (j < S && (lane % LPH) == 0 && c * 256 + lane * 8 < HD) {
          const int h = c * HPC + lane / LPH;
          sc[h * SP + j] = s * scale;
          if (BWD) dp[h * SP + j] = d;
        }
      }
    }
  }
}

template <int DH, int NCH>
__global__ void __launch_bounds__(R0_THREADS) attention_row0_fwd_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                                       const int* __restrict__ seq_start, int H, int SP,
                                                                       float scale, __nv_bfloat16* __restrict__ out0,
                                                                       const DropCfg drop) {
  constexpr int LPH = DH / 8, HPC = 256 / DH;
  extern __shared__ float r0_smem[];
  const int HD = H * DH;
  float* sc = r0_smem;                              // [H][SP]
  float* red = sc + H * SP;                         // [R0_WARPS][HD]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int seq = blockIdx.x;
  const int t0 = seq_start[seq], S = seq_start[seq + 1] - t0;
  const int ldq = 3 * HD;
  const __nv_bfloat16* qrow = qkv + (size_t)t0 * ldq;
  const __nv_bfloat16* kbase = qrow + HD;
  const __nv_bfloat16* vbase = qrow + 2 * HD;
  float q[NCH][8];
#pragma unroll
  for (int c = 0; c < NCH; ++c)
    bf16x8_to_f32(c * 256 + lane * 8 < HD ? __ldg(reinterpret_cast<const uint4*>(qrow + c * 256) + lane) : make_uint4(0, 0, 0, 0), q[c]);
  row0_scores<DH, NCH, false>(kbase, vbase, ldq, HD, S, SP, q, q, scale, sc, nullptr, warp, lane);
  __syncthreads();
  // softmax, one warp per head; the stored weight is p_j m_j (the denominator uses the full probabilities)
  const uint32_t rk = drop_row_key(drop, (uint32_t)t0);
  for (int h = warp; h < H; h += R0_WARPS) {
    float* row = sc + h * SP;
    float mx = -INFINITY;
    for (int j = lane; j < S; j += 32) mx = fmaxf(mx, row[j]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < S; j += 32) { const float e = __expf(row[j] - mx); row[j] = e; sum += e; }
    const float inv = 1.0f / warp_sum(sum);
    for (int j = lane; j < S; j += 32) row[j] *= inv * row0_mask(drop, rk, h, j);
  }
  __syncthreads();
  // o_0 = sum_j (p_j m_j) v_j: every warp sums its keys, then the warps are combined through shared memory
  float acc[NCH][8];
#pragma unroll
  for (int c = 0; c < NCH; ++c)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[c][i] = 0.f;
  for (int j0 = warp; j0 < S; j0 += R0_WARPS * R0_UN) {
    uint4 vv[R0_UN][NCH];
#pragma unroll
    for (int u = 0; u < R0_UN; ++u) {
      const int j = j0 + u * R0_WARPS;
#pragma unroll
      for (int c = 0; c < NCH; ++c)
        vv[u][c] = (j < S && c * 256 + lane * 8 < HD) ? __ldg(reinterpret_cast<const uint4*>(vbase + (size_t)j * ldq + c * 256) + lane) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < R0_UN; ++u) {
      const int j = j0 + u * R0_WARPS;
      if (j >= S) break;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        if (c * 256 + lane * 8 >= HD) continue;
        const float pj = sc[(c * HPC + lane / LPH) * SP + j];
        float f[8];
        bf16x8_to_f32(vv[u][c], f);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[c][i] = fmaf(pj, f[i], acc[c][i]);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    if (c * 256 + lane * 8 >= HD) continue;
    float4* dst = reinterpret_cast<float4*>(red + warp * HD + c * 256 + lane * 8);
    dst[0] = make_float4(acc[c][0], acc[c][1], acc[c][2], acc[c][3]);
    dst[1] = make_float4(acc[c][4], acc[c][5], acc[c][6], acc[c][7]);
  }
  __syncthreads();
  for (int d = threadIdx.x; d < HD; d += R0_THREADS) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < R0_WARPS; ++w) s += red[w * HD + d];
    out0[(size_t)seq * HD + d] = __float2bfloat16(s);
  }
}

template <int DH, int NCH>
__global__ void __launch_bounds__(R0_THREADS) attention_row0_bwd_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                                       const __nv_bfloat16* __restrict__ dout0,
                                                                       const int* __restrict__ seq_start, int H, int SP,
                                                                       float scale, __nv_bfloat16* __restrict__ dqkv,
                                                                       const DropCfg drop) {
  constexpr int LPH = DH / 8, HPC = 256 / DH;
  extern __shared__ float r0_smem[];
  const int HD = H * DH;
  float* sc = r0_smem;                              // [H][SP]  scores -> p -> ds (scale folded in)
  float* dp = sc + H * SP;                          // [H][SP]  do_0 . v_j -> p_j m_j
  float* red = dp + H * SP;                         // [R0_WARPS][HD]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int seq = blockIdx.x;
  const int t0 = seq_start[seq], S = seq_start[seq + 1] - t0;
  const int ldq = 3 * HD;
  const __nv_bfloat16* qrow = qkv + (size_t)t0 * ldq;
  const __nv_bfloat16* kbase = qrow + HD;
  const __nv_bfloat16* vbase = qrow + 2 * HD;
  float q[NCH][8], g[NCH][8];                       // this lane's columns of q_0 / do_0
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const bool act = c * 256 + lane * 8 < HD;
    bf16x8_to_f32(act ? __ldg(reinterpret_cast<const uint4*>(qrow + c * 256) + lane) : make_uint4(0, 0, 0, 0), q[c]);
    bf16x8_to_f32(act ? __ldg(reinterpret_cast<const uint4*>(dout0 + (size_t)seq * HD + c * 256) + lane) : make_uint4(0, 0, 0, 0), g[c]);
  }
  row0_scores<DH, NCH, true>(kbase, vbase, ldq, HD, S, SP, q, g, scale, sc, dp, warp, lane);
  __syncthreads();
  // per head: p = softmax(s); dp_j <- m_j dp_j (gradient w.r.t. p_j); D = sum_j p_j dp_j; ds_j = p_j (dp_j - D) * scale
  const uint32_t rk = drop_row_key(drop, (uint32_t)t0);
  for (int h = warp; h < H; h += R0_WARPS) {
    float* row = sc + h * SP;
    float* drow = dp + h * SP;
    float mx = -INFINITY;
    for (int j = lane; j < S; j += 32) mx = fmaxf(mx, row[j]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < S; j += 32) { const float e = __expf(row[j] - mx); row[j] = e; sum += e; }
    const float inv = 1.0f / warp_sum(sum);
    float dsum = 0.f;
    for (int j = lane; j < S; j += 32) {
      const float p = row[j] * inv;
      row[j] = p;
      dsum = fmaf(p, row0_mask(drop, rk, h, j) * drow[j], dsum);
    }
    const float D = warp_sum(dsum);
    for (int j = lane; j < S; j += 32) {
      const float p = row[j], m = row0_mask(drop, rk, h, j);
      row[j] = p * (m * drow[j] - D) * scale;
      drow[j] = p * m;                              // weight of v_j in o_0 = weight of do_0 in dv_j
    }
  }
  __syncthreads();
  // dk_j = ds_j q_0, dv_j = (p_j m_j) do_0, dq_{j>0} = 0 as 16-byte stores; dq_0 = sum_j ds_j k_j
  __nv_bfloat16* dq = dqkv + (size_t)t0 * ldq;
  __nv_bfloat16* dk = dq + HD;
  __nv_bfloat16* dv = dq + 2 * HD;
  float acc[NCH][8];
#pragma unroll
  for (int c = 0; c < NCH; ++c)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[c][i] = 0.f;
  for (int j0 = warp; j0 < S; j0 += R0_WARPS * R0_UN) {
    uint4 kk[R0_UN][NCH];
#pragma unroll
    for (int u = 0; u < R0_UN; ++u) {
      const int j = j0 + u * R0_WARPS;
#pragma unroll
      for (int c = 0; c < NCH; ++c)
        kk[u][c] = (j < S && c * 256 + lane * 8 < HD) ? __ldg(reinterpret_cast<const uint4*>(kbase + (size_t)j * ldq + c * 256) + lane) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < R0_UN; ++u) {
      const int j = j0 + u * R0_WARPS;
      if (j >= S) break;
      const size_t ro = (size_t)j * ldq;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        if (c * 256 + lane * 8 >= HD) continue;
        const int h = c * HPC + lane / LPH;
        const float dsj = sc[h * SP + j], pmj = dp[h * SP + j];
        float f[8];
        bf16x8_to_f32(kk[u][c], f);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[c][i] = fmaf(dsj, f[i], acc[c][i]);
        reinterpret_cast<uint4*>(dk + ro + c * 256)[lane] = scaled8(dsj, q[c]);
        reinterpret_cast<uint4*>(dv + ro + c * 256)[lane] = scaled8(pmj, g[c]);
        if (j > 0) reinterpret_cast<uint4*>(dq + ro + c * 256)[lane] = make_uint4(0, 0, 0, 0);   // query rows >= 1 carry no gradient
      }
    }
  }
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    if (c * 256 + lane * 8 >= HD) continue;
    float4* dst = reinterpret_cast<float4*>(red + warp * HD + c * 256 + lane * 8);
    dst[0] = make_float4(acc[c][0], acc[c][1], acc[c][2], acc[c][3]);
    dst[1] = make_float4(acc[c][4], acc[c][5], acc[c][6], acc[c][7]);
  }
  __syncthreads();
  for (int d = threadIdx.x; d < HD; d += R0_THREADS) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < R0_WARPS; ++w) s += red[w * HD + d];
    dq[d] = __float2bfloat16(s);
  }
}

}  // namespace eavit

using namespace eavit;

// chunks of 256 columns per section: the kernels are instantiated for 1, 2 and 4 (H * Dh = 256, 512, 1024)
#define R0_ARGS_OK()                                                                                        \
  EAVIT_CHECK_ARG(nseq > 0 && H > 0 && (Dh == 32 || Dh == 64));                                             \
  EAVIT_CHECK_ARG(H * Dh <= 256 || H * Dh == 512 || H * Dh == 1024);                                        \
  EAVIT_CHECK_ARG(max_len > 0 && max_len <= R0_MAXLEN && drop_p >= 0.f && drop_p < 1.f)

template <int DH, int NCH>
static int row0_fwd_go(const void* qkv, const int* seq_start, int nseq, int H, int SP, float scale, void* out0, const DropCfg& drop,
                       cudaStream_t st) {
  const size_t smem = ((size_t)H * SP + (size_t)R0_WARPS * H * DH) * sizeof(float);
  static size_t smem_set = 0;              // raise the limit only when a larger request appears (never during graph capture replays)
  if (smem > smem_set) {
    EAVIT_CUDA(cudaFuncSetAttribute(attention_row0_fwd_kernel<DH, NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  attention_row0_fwd_kernel<DH, NCH><<<nseq, R0_THREADS, smem, st>>>((const __nv_bfloat16*)qkv, seq_start, H, SP, scale,
                                                                      (__nv_bfloat16*)out0, drop);
  EAVIT_LAUNCH_OK();
  return EAVIT_OK;
}
template <int DH, int NCH>
static int row0_bwd_go(const void* qkv, const void* dout0, const int* seq_start, int nseq, int H, int SP, float scale, void* dqkv,
                       const DropCfg& drop, cudaStream_t st) {
  const size_t smem = ((size_t)2 * H * SP + (size_t)R0_WARPS * H * DH) * sizeof(float);
  static size_t smem_set = 0;
  if (smem > smem_set) {
    EAVIT_CUDA(cudaFuncSetAttribute(attention_row0_bwd_kernel<DH, NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  attention_row0_bwd_kernel<DH, NCH><<<nseq, R0_THREADS, smem, st>>>((const __nv_bfloat16*)qkv, (const __nv_bfloat16*)dout0,
                                                                      seq_start, H, SP, scale, (__nv_bfloat16*)dqkv, drop);
  EAVIT_LAUNCH_OK();
  return EAVIT_OK;
}

extern "C" int eavit_attention_row0_fwd(const void* qkv, const int* seq_start, int nseq, int max_len, int H, int Dh, float scale,
                                        void* out0, float drop_p, unsigned long long drop_seed, void* stream) {
  EAVIT_CHECK_ARG(qkv && seq_start && out0);
  R0_ARGS_OK();
  EAVIT_CHECK_ARG((reinterpret_cast<uintptr_t>(qkv) & 15) == 0);
  const DropCfg drop = make_drop(drop_p, drop_seed);
  const int SP = (max_len + 3) & ~3;
  cudaStream_t st = (cudaStream_t)stream;
  const int nch = (H * Dh + 255) / 256;
  if (Dh == 32) {
    if (nch == 1) return row0_fwd_go<32, 1>(qkv, seq_start, nseq, H, SP, scale, out0, drop, st);
    if (nch == 2) return row0_fwd_go<32, 2>(qkv, seq_start, nseq, H, SP, scale, out0, drop, st);
    return row0_fwd_go<32, 4>(qkv, seq_start, nseq, H, SP, scale, out0, drop, st);
  }
  if (nch == 1) return row0_fwd_go<64, 1>(qkv, seq_start, nseq, H, SP, scale, out0, drop, st);
  if (nch == 2) return row0_fwd_go<64, 2>(qkv, seq_start, nseq, H, SP, scale, out0, drop, st);
  return row0_fwd_go<64, 4>(qkv, seq_start, nseq, H, SP, scale, out0, drop, st);
}

extern "C" int eavit_attention_row0_bwd(const void* qkv, const void* dout0, const int* seq_start, int nseq, int max_len, int H,
                                        int Dh, float scale, void* dqkv, float drop_p, unsigned long long drop_seed,
                                        void* stream) {
  EAVIT_CHECK_ARG(qkv && dout0 && seq_start && dqkv);
  R0_ARGS_OK();
  EAVIT_CHECK_ARG((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(dqkv) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(dout0) & 15) == 0);
  const DropCfg drop = make_drop(drop_p, drop_seed);
  const int SP = (max_len + 3) & ~3;
  cudaStream_t st = (cudaStream_t)stream;
  const int nch = (H * Dh + 255) / 256;
  if (Dh == 32) {
    if (nch == 1) return row0_bwd_go<32, 1>(qkv, dout0, seq_start, nseq, H, SP, scale, dqkv, drop, st);
    if (nch == 2) return row0_bwd_go<32, 2>(qkv, dout0, seq_start, nseq, H, SP, scale, dqkv, drop, st);
    return row0_bwd_go<32, 4>(qkv, dout0, seq_start, nseq, H, SP, scale, dqkv, drop, st);
  }
  if (nch == 1) return row0_bwd_go<64, 1>(qkv, dout0, seq_start, nseq, H, SP, scale, dqkv, drop, st);
  if (nch == 2) return row0_bwd_go<64, 2>(qkv, dout0, seq_start, nseq, H, SP, scale, dqkv, drop, st);
  return row0_bwd_go<64, 4>(qkv, dout0, seq_start, nseq, H, SP, scale, dqkv, drop, st);
}
