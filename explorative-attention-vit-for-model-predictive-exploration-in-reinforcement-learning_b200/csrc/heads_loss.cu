// PPO policy / value heads (fp32, tiny: 0.4 MFLOP per sample) and the fused PPO + RND loss forward/backward.
//
// Reference: model.py:227-246 (heads), model.py:272-296 / :310-336 (how the two pooled features are used),
// agents.py:301-303 (old log-prob), :333-338 (masked RND loss), :455-493 (PPO-clip, value MSEs, entropy, total).
#include "common.cuh"

namespace eavit {

// C[M,N] (+)= mask(act(op(A) . op(B) + bias))      fp32, 32x32 tiles, 16x16 threads x (2x2) outputs
//   transA = 0: A[M,K] pitch lda ; 1: A stored [K,M]
//   transB = 0: B[N,K] pitch ldb (nn.Linear weight) ; 1: B stored [K,N]
// 64 x 64 output tile per CTA, 4 x 4 micro-tile per thread, K in steps of 32; blockIdx.z splits K (accumulate only,
// atomicAdd) so that the weight-gradient shapes (M = N = 256, K = rows) fill the machine instead of 16 CTAs.
__global__ void __launch_bounds__(256) sgemm_small_kernel(const float* __restrict__ A, long long lda, int transA,
                                                          const float* __restrict__ B, long long ldb, int transB,
                                                          const float* __restrict__ bias, const float* __restrict__ mask_aux,
                                                          float* __restrict__ C, long long ldc, int M, int N, int K, int relu,
                                                          int accumulate, int k_per_split) {
  // These GEMMs are latency-bound (a handful of CTAs, K = 256 .. 1024): K advances 32 at a time and the next step's
  // operands are fetched into registers (16 independent loads per thread) while the current step is multiplied.
  constexpr int SK = 32, LD = 65;        // odd pitch: conflict-free for both the k-fastest and the m-fastest fill
  __shared__ float sA[SK][LD];   // [k][m]
  __shared__ float sB[SK][LD];   // [k][n]
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int kbeg = blockIdx.z * k_per_split, kend = min(K, kbeg + k_per_split);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float ra[8], rb[8];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int i = threadIdx.x + 256 * r;
      int kk, mm;
      if (transA) { mm = i & 63; kk = i >> 6; } else { kk = i & 31; mm = i >> 5; }
      const int gm = m0 + mm, gk = k0 + kk;
      ra[r] = (gm < M && gk < kend) ? __ldg(transA ? A + (size_t)gk * lda + gm : A + (size_t)gm * lda + gk) : 0.f;
      int nn;
      if (transB) { nn = i & 63; kk = i >> 6; } else { kk = i & 31; nn = i >> 5; }
      const int gn = n0 + nn, gk2 = k0 + kk;
      rb[r] = (gn < N && gk2 < kend) ? __ldg(transB ? B + (size_t)gk2 * ldb + gn : B + (size_t)gn * ldb + gk2) : 0.f;
    }
  };
  fetch(kbeg);
  for (int k0 = kbeg; k0 < kend; k0 += SK) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int i = threadIdx.x + 256 * r;
      if (transA) sA[i >> 6][i & 63] = ra[r]; else sA[i & 31][i >> 5] = ra[r];
      if (transB) sB[i >> 6][i & 63] = rb[r]; else sB[i & 31][i >> 5] = rb[r];
    }
    __syncthreads();
    if (k0 + SK < kend) fetch(k0 + SK);
#pragma unroll
    for (int kk = 0; kk < SK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = sA[kk][ty + 16 * i]; b[i] = sB[kk][tx + 16 * i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gm = m0 + ty + 16 * i, gn = n0 + tx + 16 * j;
      if (gm < M && gn < N) {
        float v = acc[i][j];
        float* c = C + (size_t)gm * ldc + gn;
        if (gridDim.z > 1) { atomicAdd(c, v); continue; }      // split-K: plain accumulation only (checked on the host)
        if (bias != nullptr) v += bias[gn];
        if (relu) v = fmaxf(v, 0.f);
        if (mask_aux != nullptr && !(mask_aux[(size_t)gm * ldc + gn] > 0.f)) v = 0.f;
        *c = accumulate ? (*c + v) : v;
      }
    }
}

// values: rows r in [0, R): h = E[r] + F[r]; v[r] = h . w(r) + b(r), w = (r < split ? wA : wB)
__global__ void __launch_bounds__(256) heads_value_fwd_kernel(const float* __restrict__ E, const float* __restrict__ F,
                                                              const float* __restrict__ wA, const float* __restrict__ bA,
                                                              const float* __restrict__ wB, const float* __restrict__ bB,
                                                              int split, int R, int D, float* __restrict__ v) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= R) return;
  const float* w = r < split ? wA : wB;
  float a = 0.f;
  for (int c = lane; c < D; c += 32) a += (E[(size_t)r * D + c] + F[(size_t)r * D + c]) * w[c];
  a = warp_sum(a);
  if (lane == 0) v[r] = a + (r < split ? bA[0] : bB[0]);
}
// backward of the above: dH[r] = dv[r] * w(r)  -> dF[r] = dH, dE[r] = dH * (E > 0);  dw += sum dv[r] h[r]; db += sum dv
// single block per weight-half keeps the dw reduction deterministic (R <= a few thousand rows, D = 256).
__global__ void __launch_bounds__(256) heads_value_bwd_kernel(const float* __restrict__ E, const float* __restrict__ F,
                                                              const float* __restrict__ dv, const float* __restrict__ wA,
                                                              const float* __restrict__ wB, int split, int R, int D,
                                                              float* __restrict__ dE, float* __restrict__ dF,
                                                              float* __restrict__ dwA, float* __restrict__ dbA,
                                                              float* __restrict__ dwB, float* __restrict__ dbB) {
  // one CTA per chunk of 32 rows (was: 2 CTAs looping over R/2 rows each -- 125 us of pure latency); the weight / bias
  // gradients of the two critics are combined with one atomicAdd per column per CTA.  Rows [0,split) use head A.
  const int r0 = blockIdx.x * 32, r1 = min(R, r0 + 32);
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    const float wa = wA[c], wb = wB[c];
    float accA = 0.f, accB = 0.f;
    for (int r = r0; r < r1; ++r) {
      const float e = E[(size_t)r * D + c], f = F[(size_t)r * D + c], g = dv[r];
      const bool isA = r < split;
      const float dh = g * (isA ? wa : wb);
      dF[(size_t)r * D + c] = dh;
      dE[(size_t)r * D + c] = e > 0.f ? dh : 0.f;
      if (isA) accA += g * (e + f); else accB += g * (e + f);
    }
    if (r0 < split) atomicAdd(dwA + c, accA);
    if (r1 > split) atomicAdd(dwB + c, accB);
  }
  if (threadIdx.x == 0) {
    float sa = 0.f, sb = 0.f;
    for (int r = r0; r < r1; ++r) { if (r < split) sa += dv[r]; else sb += dv[r]; }
    if (r0 < split) atomicAdd(dbA, sa);
    if (r1 > split) atomicAdd(dbB, sb);
  }
}

// comb[b] = 0.5 (F[b] + F[B+b])  (model.py:284-286 'mean') or sum ('sum', model.py:287-288)
__global__ void combine_fwd_kernel(const float* __restrict__ F, float* __restrict__ comb, int B, int D, float coef) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (long long)B * D) comb[i] = coef * (F[i] + F[(long long)B * D + i]);
}
// dF[b] += coef dcomb[b]; dF[B+b] += coef dcomb[b]
__global__ void combine_bwd_kernel(const float* __restrict__ dcomb, float* __restrict__ dF, int B, int D, float coef) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (long long)B * D) { const float g = coef * dcomb[i]; dF[i] += g; dF[(long long)B * D + i] += g; }
}
__global__ void add_f32_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] + b[i];
}

// ----------------------------------------------------------------------------------------------
// PPO loss, forward + backward in one pass over [B, A] logits (A <= 32).  One block.
// stats: [1]=actor [2]=critic_ext [3]=critic_int [4]=entropy [6]=approx_kl [7]=max_kl [8]=clipfrac
// ----------------------------------------------------------------------------------------------
constexpr int PPO_MAXA = 32;
__global__ void __launch_bounds__(1024) ppo_loss_kernel(const float* __restrict__ logits, const float* __restrict__ old_logits,
                                                        const long long* __restrict__ actions, const float* __restrict__ adv,
                                                        const float* __restrict__ v_ext, const float* __restrict__ v_int,
                                                        const float* __restrict__ tgt_ext, const float* __restrict__ tgt_int,
                                                        int B, int A, float ppo_eps, float ent_coef, float grad_scale,
                                                        float* __restrict__ dlogits, float* __restrict__ dv_ext,
                                                        float* __restrict__ dv_int, float* __restrict__ stats) {
  float s_actor = 0.f, s_ce = 0.f, s_ci = 0.f, s_ent = 0.f, s_kl = 0.f, s_clip = 0.f, m_kl = -INFINITY;
  const float invB = 1.f / (float)B;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    float z[PPO_MAXA], zo[PPO_MAXA];
    float mx = -INFINITY, mxo = -INFINITY;
#pragma unroll
    for (int k = 0; k < PPO_MAXA; ++k)
      if (k < A) { z[k] = logits[(size_t)b * A + k]; zo[k] = old_logits[(size_t)b * A + k]; mx = fmaxf(mx, z[k]); mxo = fmaxf(mxo, zo[k]); }
    float se = 0.f, seo = 0.f;
#pragma unroll
    for (int k = 0; k < PPO_MAXA; ++k)
      if (k < A) { se += expf(z[k] - mx); seo += expf(zo[k] - mxo); }
    const float lse = mx + logf(se), lseo = mxo + logf(seo);
    const int y = (int)actions[b];
    float logp_y = 0.f, logpo_y = 0.f, H = 0.f;
#pragma unroll
    for (int k = 0; k < PPO_MAXA; ++k)
      if (k < A) {
        const float lp = z[k] - lse;
        H -= expf(lp) * lp;
        if (k == y) { logp_y = lp; logpo_y = zo[k] - lseo; }
      }
    const float ratio = expf(logp_y - logpo_y);                                   // agents.py:466
    const float a = adv[b];
    const float lo = 1.f - ppo_eps, hi = 1.f + ppo_eps;
    const float rc = fminf(fmaxf(ratio, lo), hi);
    const float surr1 = ratio * a, surr2 = rc * a;                                // agents.py:468-472
    s_actor += -fminf(surr1, surr2);
    // d min(surr1, surr2) / d ratio  (torch.min splits ties evenly; clamp passes gradient on [lo, hi])
    const float in_rng = (ratio >= lo && ratio <= hi) ? 1.f : 0.f;
    float dr;
    if (surr1 < surr2) dr = a; else if (surr2 < surr1) dr = a * in_rng; else dr = 0.5f * a + 0.5f * a * in_rng;
    const float dlogp = -dr * ratio * invB;                                       // d actor_loss / d log_prob
    const float ve = v_ext[b], vi = v_int[b];
    const float de = ve - tgt_ext[b], di = vi - tgt_int[b];
    s_ce += de * de; s_ci += di * di;                                             // F.mse_loss, agents.py:476-479
    dv_ext[b] = grad_scale * de * invB;                                           // 0.5 * d mean((v-R)^2)
    dv_int[b] = grad_scale * di * invB;
    s_ent += H;
    const float dkl = logpo_y - logp_y;
    s_kl += dkl; m_kl = fmaxf(m_kl, dkl);
    s_clip += (ratio > hi || ratio < lo) ? 1.f : 0.f;
#pragma unroll
    for (int k = 0; k < PPO_MAXA; ++k)
      if (k < A) {
        const float lp = z[k] - lse, p = expf(lp);
        float g = dlogp * ((k == y ? 1.f : 0.f) - p);                             // d log_softmax
        g += ent_coef * invB * p * (lp + H);                                      // d(-ent_coef * mean H)
        dlogits[(size_t)b * A + k] = grad_scale * g;
      }
  }
  __shared__ float red[7][32];
  float vals[6] = {s_actor, s_ce, s_ci, s_ent, s_kl, s_clip};
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < 6; ++i) { const float t = warp_sum(vals[i]); if (lane == 0) red[i][w] = t; }
  { const float t = warp_max(m_kl); if (lane == 0) red[6][w] = t; }
  __syncthreads();
  if (w == 0) {
    const int nw = blockDim.x >> 5;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      float t = lane < nw ? red[i][lane] : 0.f;
      t = warp_sum(t);
      vals[i] = t;
    }
    float t = lane < nw ? red[6][lane] : -INFINITY;
    t = warp_max(t);
    if (lane == 0) {
      stats[1] = vals[0] * invB; stats[2] = vals[1] * invB; stats[3] = vals[2] * invB; stats[4] = vals[3] * invB;
      stats[6] = vals[4] * invB; stats[7] = t; stats[8] = vals[5] * invB;
    }
  }
}

// RND loss (agents.py:333-338): per = mean_j (pred - tgt)^2 ; loss = sum(per * mask) / max(sum mask, 1)
// dpred[i,j] = grad_scale * 2 (pred - tgt) / R * mask_i / max(sum mask, 1).   stats[5] += loss contribution.
__global__ void __launch_bounds__(256) rnd_loss_kernel(const float* __restrict__ pred, const float* __restrict__ tgt,
                                                       const float* __restrict__ mask, int B, int R, float grad_scale,
                                                       __nv_bfloat16* __restrict__ dpred_bf16, float* __restrict__ per_out,
                                                       float* __restrict__ stats) {
  __shared__ float s_msum;
  __shared__ float s_red[8];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float ms = 0.f;
  for (int i = threadIdx.x; i < B; i += blockDim.x) ms += mask[i];
  ms = warp_sum(ms);
  if (lane == 0) s_red[w] = ms;
  __syncthreads();
  if (threadIdx.x == 0) { float t = 0.f; for (int i = 0; i < 8; ++i) t += s_red[i]; s_msum = fmaxf(t, 1.f); }
  __syncthreads();
  const float denom = s_msum;
  const int row = blockIdx.x * 8 + w;
  float contrib = 0.f;
  if (row < B) {
    const float m = mask[row];
    const float gs = grad_scale * 2.f / (float)R * m / denom;
    float acc = 0.f;
    for (int c = lane; c < R / 2; c += 32) {
      const float2 p = *reinterpret_cast<const float2*>(pred + (size_t)row * R + 2 * c);
      const float2 t = *reinterpret_cast<const float2*>(tgt + (size_t)row * R + 2 * c);
      const float d0 = p.x - t.x, d1 = p.y - t.y;
      acc += d0 * d0 + d1 * d1;
      if (dpred_bf16 != nullptr)
        *reinterpret_cast<uint32_t*>(dpred_bf16 + (size_t)row * R + 2 * c) = pack_bf16x2(gs * d0, gs * d1);
    }
    acc = warp_sum(acc) / (float)R;
    if (lane == 0 && per_out != nullptr) per_out[row] = acc;
    contrib = acc * m / denom;
  }
  __syncthreads();
  if (lane == 0) s_red[w] = contrib;
  __syncthreads();
  if (threadIdx.x == 0) { float t = 0.f; for (int i = 0; i < 8; ++i) t += s_red[i]; atomicAdd(stats + 5, t); }
}

__global__ void gather_batch_kernel(const long long* __restrict__ idx, int B, int A, const float* __restrict__ te,
                                    const float* __restrict__ ti, const float* __restrict__ adv,
                                    const long long* __restrict__ act, const float* __restrict__ old,
                                    float* __restrict__ o_te, float* __restrict__ o_ti, float* __restrict__ o_adv,
                                    long long* __restrict__ o_act, float* __restrict__ o_old) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * (A + 1)) return;
  const int b = i / (A + 1), k = i % (A + 1);
  const long long s = idx[b];
  if (k == A) { o_te[b] = te[s]; o_ti[b] = ti[s]; o_adv[b] = adv[s]; o_act[b] = act[s]; }
  else o_old[(size_t)b * A + k] = old[(size_t)s * A + k];
}

}  // namespace eavit

using namespace eavit;

extern "C" {

int eavit_sgemm_small(const float* A, long long lda, int transA, const float* B, long long ldb, int transB,
                      const float* bias, const float* mask_aux, float* C, long long ldc, int M, int N, int K, int relu,
                      int accumulate, void* stream) {
  EAVIT_CHECK_ARG(A && B && C && M > 0 && N > 0 && K > 0);
  // split K when the output has too few tiles to fill the GPU (weight gradients: K = number of rows) -- accumulate only
  const int tiles = cdiv(N, 64) * cdiv(M, 64);
  int splits = 1;
  if (accumulate && !bias && !mask_aux && !relu && tiles < kNumSMs && K >= 256) {
    splits = kNumSMs / tiles;
    if (splits > K / 64) splits = K / 64;
    if (splits < 1) splits = 1;
  }
  int kps = cdiv(cdiv(K, splits), 32) * 32;
  splits = cdiv(K, kps);
  dim3 grid(cdiv(N, 64), cdiv(M, 64), splits);
  sgemm_small_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(A, lda, transA, B, ldb, transB, bias, mask_aux, C, ldc, M, N, K, relu, accumulate, kps);
  EAVIT_LAUNCH_OK();
  return EAVIT_OK;
}

int eavit_heads_value_fwd(const float* E, const float* F, const float* wA, const float* bA, const float* wB,
                          const float* bB, int split, int R, int D, float* v, void* stream) {
  EAVIT_CHECK_ARG(E && F && wA && bA && wB && bB && v && R > 0 && D > 0 && split >= 0 && split <= R);
  heads_value_fwd_kernel<<<cdiv(R, 8), 256, 0, (cudaStream_t)stream>>>(E, F, wA, bA, wB, bB, split, R, D, v);
  EAVIT_LAUNCH_OK();
  return EAVIT_OK;
}

int eavit_heads_value_bwd(const float* E, const float* F, const float* dv, const float* wA, const float* wB, int split,
                          int R, int D, float* dE, float* dF, float* dwA, float* dbA, float* dwB, float* dbB, void* stream) {
  EAVIT_CHECK_ARG(E && F && dv && wA && wB && dE && dF && dwA && dbA && dwB && dbB && R > 0 && D > 0);
  heads_value_bwd_kernel<<<cdiv(R, 32), 256, 0, (cudaStream_t)stream>>>(E, F, dv, wA, wB, split, R, D, dE, dF, dwA, dbA, dwB, dbB);
  EAVIT_LAUNCH_OK();
  return EAVIT_OK;
}

int eavit_combine_fwd(const float* F, float* comb, int B, int D, float coef, void* stream) {
  EAVIT_CHECK_ARG(F && comb && B > 0 && D > 0);
  combine_fwd_kernel<<<cdiv((long long)B * D, 256), 256, 0, (cudaStream_t)stream>>>(F, comb, B, D, coef);
  EAVIT_LAUNCH_OK();
  return EAVIT_OK;
}
int eavit_combine_bwd(const float* dcomb, float* dF, int B, int D, float coef, void* stream) {
  EAVIT_CHECK_ARG(dcomb && dF && B > 0 && D > 0);
  combine_bwd_kernel<<<cdiv((long long)B * D, 256), 256, 0, (cudaStream_t)stream>>>(dcomb, dF, B, D, coef);
  EAVIT_LAUNCH_OK();
  return EAVIT_OK;
}
int eavit_add_f32(const float* a, const float* b, float* out, long long n, void* stream) {
  EAVIT_CHECK_ARG(a && b && out && n > 0);
  add_f32_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(a, b, out, n);
  EAVIT_LAUNCH_OK();
  return EAVIT_OK;
}

int eavit_ppo_loss(const float* logits, const float* old_logits, const long long* actions, const float* adv,
                   const float* v_ext, const float* v_int, const float* tgt_ext, const float* tgt_int, int B, int A,
                   float ppo_eps, float ent_coef, float grad_scale, float* dlogits, float* dv_ext, float* dv_int,
                   float* stats, void* stream) {
  EAVIT_CHECK_ARG(logits && old_logits && actions && adv && v_ext && v_int && tgt_ext && tgt_int && dlogits && dv_ext && dv_int && stats);
  EAVIT_CHECK_ARG(B > 0 && A > 0 && A <= PPO_MAXA);
  int threads = ((B + 31) / 32) * 32;
  if (threads > 1024) threads = 1024;
  ppo_loss_kernel<<<1, threads, 0, (cudaStream_t)stream>>>(logits, old_logits, actions, adv, v_ext, v_int, tgt_ext, tgt_int, B, A,
                                                          ppo_eps, ent_coef, grad_scale, dlogits, dv_ext, dv_int, stats);
  EAVIT_LAUNCH_OK();
  return EAVIT_OK;
}

int eavit_rnd_loss(const float* pred, const float* tgt, const float* mask, int B, int R, float grad_scale, void* dpred_bf16,
                   float* per_sample, float* stats, void* stream) {
  EAVIT_CHECK_ARG(pred && tgt && mask && stats && B > 0 && R > 0 && R % 2 == 0);
  rnd_loss_kernel<<<cdiv(B, 8), 256, 0, (cudaStream_t)stream>>>(pred, tgt, mask, B, R, grad_scale, (__nv_bfloat16*)dpred_bf16, per_sample, stats);
  EAVIT_LAUNCH_OK();
  return EAVIT_OK;
}

int eavit_gather_batch(const long long* idx, int B, int A, const float* tgt_ext, const float* tgt_int, const float* adv,
                       const long long* actions, const float* old_logits, float* o_tgt_ext, float* o_tgt_int, float* o_adv,
                       long long* o_actions, float* o_old_logits, void* stream) {
  EAVIT_CHECK_ARG(idx && tgt_ext && tgt_int && adv && actions && old_logits && o_tgt_ext && o_tgt_int && o_adv && o_actions && o_old_logits);
  EAVIT_CHECK_ARG(B > 0 && A > 0);
  gather_batch_kernel<<<cdiv((long long)B * (A + 1), 256), 256, 0, (cudaStream_t)stream>>>(idx, B, A, tgt_ext, tgt_int, adv, actions, old_logits,
                                                                                  o_tgt_ext, o_tgt_int, o_adv, o_actions, o_old_logits);
  EAVIT_LAUNCH_OK();
  return EAVIT_OK;
}

int eavit_zero(void* ptr, long long bytes, void* stream) {
  EAVIT_CHECK_ARG(ptr && bytes >= 0);
  EAVIT_CUDA(cudaMemsetAsync(ptr, 0, (size_t)bytes, (cudaStream_t)stream));
  return EAVIT_OK;
}

}  // extern "C"
