// Patch embedding front-end and token/positional assembly (and their backward).
//
//   patchify        : [B,C,H,W] (u8 -> /255, or f32) -> bf16 [B*np, PD]; optional fused LayerNorm(PD)
//                     lucidrains order (p1 p2 c), vit.py:110-111;  HF conv order (c p1 p2), ViTPatchEmbeddings
//   embed_assemble  : token prepend + positional add into the flat fp32 residual stream, reproducing the
//                     reference's token bug (vit.py:141-156) in mode 0
//   *_bwd           : gradients of the above (pos_embedding, tokens, LayerNorm(PD) affine)
//
// One CTA stages one patch-row of one sample (C x p x W pixels, coalesced) in shared memory; warps then
// walk the patches of that row.  An optional sample index (minibatch gather from the device-resident
// rollout) replaces the reference's per-minibatch host re-materialisation (agents.py:288).
#include "common.cuh"

namespace eavit {

constexpr int PATCH_MAXV = 18;   // PD <= 576 = 32 * 18

template <typename ImgT>
__device__ __forceinline__ float load_pixel(const ImgT* p) {
  if constexpr (sizeof(ImgT) == 1) return __fdiv_rn((float)(*p), 255.0f);   // np.float32(states) / 255.  (train.py:605,854)
  else return *p;
}

// MODE 0: forward (write bf16 patches [+LN]);  MODE 1: backward of the LN affine (dgamma, dbeta).
// VPL = ceil(PD / 32) patch elements per lane (5 for the 6x6x4 patches, 18 for 12x12x4).  The (p1 p2 c) / (c p1 p2)
// element order is turned into a per-CTA offset table once, so the per-element work has no integer division (the
// first version spent most of its 0.2-0.3 ms on index arithmetic for 14 MB of pixels).
template <typename ImgT, int MODE, int VPL>
__global__ void __launch_bounds__(128) patchify_kernel(const ImgT* __restrict__ img, const long long* __restrict__ sample_idx,
                                                       int B, int C, int HW, int P, int order_cpp,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       float eps, __nv_bfloat16* __restrict__ out,
                                                       float* __restrict__ mean_io, float* __restrict__ rstd_io,
                                                       const float* __restrict__ dpln, float* __restrict__ dgamma,
                                                       float* __restrict__ dbeta) {
  extern __shared__ float tile[];     // [C][P][HW] | int koff[PD] | (MODE 1: [4 warps][2][PD])
  const int npr = HW / P;             // patches per row
  const int PD = C * P * P;
  int* koff = reinterpret_cast<int*>(tile + C * P * HW);
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int k = threadIdx.x; k < PD; k += blockDim.x) {
    int c, p1, p2;
    if (order_cpp) { c = k / (P * P); p1 = (k / P) % P; p2 = k % P; }      // (c p1 p2)
    else           { c = k % C; p2 = (k / C) % P; p1 = k / (C * P); }       // (p1 p2 c)
    koff[k] = (c * P + p1) * HW + p2;
  }
  float ag[VPL], ab[VPL], gm[VPL], bt[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int k = lane + 32 * i;
    ag[i] = 0.f; ab[i] = 0.f;
    gm[i] = (gamma != nullptr && k < PD) ? gamma[k] : 0.f;
    bt[i] = (MODE == 0 && beta != nullptr && k < PD) ? beta[k] : 0.f;
  }
  // persistent over (sample, patch-row) items: the MODE 1 atomics are paid once per CTA
  for (int item = blockIdx.x; item < B * npr; item += gridDim.x) {
    const int b = item / npr, ph = item - b * npr;
    const long long src = sample_idx ? sample_idx[b] : (long long)b;
    const ImgT* base = img + (size_t)src * C * HW * HW + (size_t)ph * P * HW;
    __syncthreads();                    // previous item's tile fully consumed (and koff written)
    for (int rr = w; rr < C * P; rr += 4) {            // one image row segment (HW contiguous pixels) per warp pass
      const int c = rr / P, r = rr - c * P;
      const ImgT* srow = base + ((size_t)c * HW + r) * HW;
      for (int x = lane; x < HW; x += 32) tile[rr * HW + x] = load_pixel(srow + x);
    }
    __syncthreads();
    for (int pw = w; pw < npr; pw += 4) {
      const size_t row = (size_t)b * npr * npr + (size_t)ph * npr + pw;
      const float* tp = tile + pw * P;
      float v[VPL];
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const int k = lane + 32 * i;
        v[i] = (k < PD) ? tp[koff[k]] : 0.f;
        s += v[i];
      }
      if (gamma == nullptr) {            // no LayerNorm: plain bf16 patches
        if (MODE == 0) {
#pragma unroll
          for (int i = 0; i < VPL; ++i) {
            const int k = lane + 32 * i;
            if (k < PD) out[row * PD + k] = __float2bfloat16(v[i]);
          }
        }
        continue;
      }
      float mu, rs;
      if (MODE == 0) {
        mu = warp_sum(s) / (float)PD;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
          const int k = lane + 32 * i;
          if (k < PD) { const float d = v[i] - mu; q += d * d; }
        }
        rs = rsqrtf(warp_sum(q) / (float)PD + eps);
        if (lane == 0 && mean_io != nullptr) { mean_io[row] = mu; rstd_io[row] = rs; }
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
          const int k = lane + 32 * i;
          if (k < PD) out[row * PD + k] = __float2bfloat16((v[i] - mu) * rs * gm[i] + bt[i]);
        }
      } else {
        mu = mean_io[row]; rs = rstd_io[row];
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
          const int k = lane + 32 * i;
          if (k < PD) {
            const float d = __ldg(dpln + row * PD + k);
            ag[i] += d * (v[i] - mu) * rs;
            ab[i] += d;
          }
        }
      }
    }
  }
  if (MODE == 1) {
    float* part = tile + C * P * HW + PD;    // [4][2][PD]
    __syncthreads();
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int k = lane + 32 * i;
      if (k < PD) { part[(w * 2) * PD + k] = ag[i]; part[(w * 2 + 1) * PD + k] = ab[i]; }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < PD; k += blockDim.x) {
      float sg = 0.f, sb = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) { sg += part[(q * 2) * PD + k]; sb += part[(q * 2 + 1) * PD + k]; }
      atomicAdd(dgamma + k, sg);
      atomicAdd(dbeta + k, sb);
    }
  }
}

// ----------------------------------------------------------------------------------------------
// assemble: e [B*np, D] fp32 (patch tokens) -> flat residual stream x [Ttot, D] fp32
//   mode 0 (lucidrains explorative pair): seqA_b = e_b                      (len np,   rows [0, B*np))
//                                         seqB_b = [tokA+pos0, e_b+pos_1..] (len np+1, rows after)
//   mode 1 (single, CLS):                 seq_b  = [tokA+pos0, e_b+pos_1..]
//   mode 2 (HF pair):                     seqA_b = [tokA+pos0, e_b+pos..], seqB_b = [tokB+pos0, e_b+pos..]
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) embed_assemble_kernel(const float* __restrict__ e, const float* __restrict__ pos,
                                                             const float* __restrict__ tokA, const float* __restrict__ tokB,
                                                             int mode, int B, int np, int D, float* __restrict__ x) {
  // one warp per (b, j) with j in [0, np]  (j = 0 is the token row)
  const int wid = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (wid >= B * (np + 1)) return;
  const int b = wid / (np + 1), j = wid % (np + 1);
  const int S1 = np + 1;
  const int nv = D / 4;
  const float4* pj = reinterpret_cast<const float4*>(pos + (size_t)j * D);
  if (j == 0) {
    for (int c = lane; c < nv; c += 32) {
      const float4 p = __ldg(pj + c);
      const float4 ta = __ldg(reinterpret_cast<const float4*>(tokA) + c);
      const float4 va = make_float4(ta.x + p.x, ta.y + p.y, ta.z + p.z, ta.w + p.w);
      if (mode == 0) {
        *(reinterpret_cast<float4*>(x + ((size_t)B * np + (size_t)b * S1) * D) + c) = va;
      } else if (mode == 1) {
        *(reinterpret_cast<float4*>(x + ((size_t)b * S1) * D) + c) = va;
      } else {
        const float4 tb = __ldg(reinterpret_cast<const float4*>(tokB) + c);
        *(reinterpret_cast<float4*>(x + ((size_t)b * S1) * D) + c) = va;
        *(reinterpret_cast<float4*>(x + ((size_t)B * S1 + (size_t)b * S1) * D) + c) =
            make_float4(tb.x + p.x, tb.y + p.y, tb.z + p.z, tb.w + p.w);
      }
    }
    return;
  }
  const int n = j - 1;
  const float4* er = reinterpret_cast<const float4*>(e + ((size_t)b * np + n) * D);
  for (int c = lane; c < nv; c += 32) {
    const float4 v = __ldg(er + c);
    const float4 p = __ldg(pj + c);
    const float4 vp = make_float4(v.x + p.x, v.y + p.y, v.z + p.z, v.w + p.w);
    if (mode == 0) {
      *(reinterpret_cast<float4*>(x + ((size_t)b * np + n) * D) + c) = v;                       // no pos-emb (bug kept)
      *(reinterpret_cast<float4*>(x + ((size_t)B * np + (size_t)b * S1 + j) * D) + c) = vp;
    } else if (mode == 1) {
      *(reinterpret_cast<float4*>(x + ((size_t)b * S1 + j) * D) + c) = vp;
    } else {
      *(reinterpret_cast<float4*>(x + ((size_t)b * S1 + j) * D) + c) = vp;
      *(reinterpret_cast<float4*>(x + ((size_t)B * S1 + (size_t)b * S1 + j) * D) + c) = vp;
    }
  }
}

// g[b*np+n] = sum over sequences of the patch-token gradient (fp32 + optional bf16 copy)
__global__ void __launch_bounds__(256) embed_assemble_bwd_rows_kernel(const float* __restrict__ dx, int mode, int B, int np,
                                                                      int D, float* __restrict__ g,
                                                                      __nv_bfloat16* __restrict__ g_bf16) {
  const int wid = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (wid >= B * np) return;
  const int b = wid / np, n = wid % np, S1 = np + 1, nv = D / 4;
  const float4 *a, *c2 = nullptr;
  if (mode == 0) {
    a = reinterpret_cast<const float4*>(dx + ((size_t)b * np + n) * D);
    c2 = reinterpret_cast<const float4*>(dx + ((size_t)B * np + (size_t)b * S1 + 1 + n) * D);
  } else if (mode == 1) {
    a = reinterpret_cast<const float4*>(dx + ((size_t)b * S1 + 1 + n) * D);
  } else {
    a = reinterpret_cast<const float4*>(dx + ((size_t)b * S1 + 1 + n) * D);
    c2 = reinterpret_cast<const float4*>(dx + ((size_t)B * S1 + (size_t)b * S1 + 1 + n) * D);
  }
  for (int c = lane; c < nv; c += 32) {
    float4 v = __ldg(a + c);
    if (c2 != nullptr) { const float4 u = __ldg(c2 + c); v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w; }
    if (g != nullptr) *(reinterpret_cast<float4*>(g + (size_t)wid * D) + c) = v;
    if (g_bf16 != nullptr)
      *(reinterpret_cast<uint2*>(g_bf16 + (size_t)wid * D) + c) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  }
}

// dpos[j] += sum_b (...), dtokA/B += sum_b dx[token row].  grid = (np+1, ceil(D/128)), block 128, loops over b.
__global__ void __launch_bounds__(128) embed_assemble_bwd_pos_kernel(const float* __restrict__ dx, int mode, int B, int np,
                                                                     int D, float* __restrict__ dpos,
                                                                     float* __restrict__ dtokA, float* __restrict__ dtokB) {
  const int j = blockIdx.x, c = blockIdx.y * 128 + threadIdx.x;
  if (c >= D) return;
  const int S1 = np + 1;
  float sa = 0.f, sb = 0.f;
  for (int b = 0; b < B; ++b) {
    if (mode == 0) {
      sb += dx[((size_t)B * np + (size_t)b * S1 + j) * D + c];
    } else if (mode == 1) {
      sa += dx[((size_t)b * S1 + j) * D + c];
    } else {
      sa += dx[((size_t)b * S1 + j) * D + c];
      sb += dx[((size_t)B * S1 + (size_t)b * S1 + j) * D + c];
    }
  }
  atomicAdd(dpos + (size_t)j * D + c, sa + sb);
  if (j == 0) {
    if (mode == 0) atomicAdd(dtokA + c, sb);          // exploration_token feeds the exploitative pass (bug kept)
    else if (mode == 1) atomicAdd(dtokA + c, sa);
    else { atomicAdd(dtokA + c, sa); atomicAdd(dtokB + c, sb); }
  }
}

// Both of the above in ONE pass over dx (D = 128 * NV4 <= 1024): a persistent grid walks (sequence position j, chunk of samples)
// items, a CTA's 8 warps take the samples in turn.  For j >= 1 a warp forms the patch-token gradient g[b, j-1] (sum over the
// sequences) and keeps the row it has just read as its share of dpos[j]; the j = 0 items only sum the token rows (dpos[0], dtokA /
// dtokB).  The 8 warps' partial sums meet in shared memory and leave as one atomicAdd per (j, column) and item -- the position /
// token gradients no longer re-read the 103 MB the separate kernel streamed with one 512-step strided loop per thread.
//   LN = true additionally runs the backward of the LayerNorm(D) that ends to_patch_embedding (vit.py:113) on the row the
// warp holds: g never goes to memory, de = LN'(g) leaves as bf16 (the dY operand of the patch Linear's dW / dX GEMMs), and
// dgamma / dbeta / the Linear's bias gradient (column sums of de) join the per-CTA partial sums.
struct EmbedLnBwd {
  const float* e0;      // LayerNorm input [B*np, D]
  const float* mean;    // [B*np]
  const float* rstd;
  const float* gamma;   // [D]
  float* dgamma;        // += sum g * xhat
  float* dbeta;         // += sum g
  float* dbias;         // += sum de   (may be NULL)
};
//   L1 = true puts one more LayerNorm backward IN FRONT of the pass: the first transformer layer's pre-attention LayerNorm
// (vit.py:47 of layer 0).  `dx` then is the residual gradient arriving at that LayerNorm's input (dres), and the gradient of
// the embedding output is formed per row as dres + LN1'(dy) from dy (bf16, the QKV dX GEMM's output), the LayerNorm's input x
// (the embedding output itself) and its statistics -- the [T, D] fp32 gradient never makes the round trip through memory.
struct EmbedLn1Bwd {
  const __nv_bfloat16* dy;   // [T, D]
  const float* x;            // [T, D]
  const float* mean;         // [T]
  const float* rstd;
  const float* gamma;        // [D]
  float* dgamma;
  float* dbeta;
};
template <int NV4, bool LN, bool L1>
__global__ void __launch_bounds__(256, (LN && NV4 <= 2) ? 2 : 1) embed_assemble_bwd_fused_kernel(const float* __restrict__ dx, int mode, int B, int np,
                                                                       float* __restrict__ g, __nv_bfloat16* __restrict__ g_bf16,
                                                                       float* __restrict__ dpos, float* __restrict__ dtokA,
                                                                       float* __restrict__ dtokB, int bchunk, int nchunk,
                                                                       const EmbedLnBwd ln, const EmbedLn1Bwd l1, const DropCfg drop) {
  constexpr int D = 128 * NV4;
  __shared__ float4 red[8][32 * NV4];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S1 = np + 1;
  // row of (sequence set q, sample b, position jj) in the flat token matrix; set 1 exists for modes 0 and 2
  auto row = [&](int q, int b, int jj) -> size_t {
    if (mode == 0) return q == 0 ? (size_t)b * np + (jj - 1) : (size_t)B * np + (size_t)b * S1 + jj;
    return (size_t)q * B * S1 + (size_t)b * S1 + jj;
  };
  auto zero = [&](float4* acc) {
#pragma unroll
    for (int i = 0; i < NV4; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  };
  // partial sums that live across every item of this (persistent) CTA: one flush per CTA at the end
  float4 ag[LN ? NV4 : 1], ab[LN ? NV4 : 1], ax[LN ? NV4 : 1], gm[LN ? NV4 : 1];
  float4 ag1[L1 ? NV4 : 1], ab1[L1 ? NV4 : 1], gm1[L1 ? NV4 : 1];
  if constexpr (LN) {
    zero(ag); zero(ab); zero(ax);
#pragma unroll
    for (int i = 0; i < NV4; ++i) gm[i] = __ldg(reinterpret_cast<const float4*>(ln.gamma) + lane + 32 * i);
  }
  if constexpr (L1) {
    zero(ag1); zero(ab1);
#pragma unroll
    for (int i = 0; i < NV4; ++i) gm1[i] = __ldg(reinterpret_cast<const float4*>(l1.gamma) + lane + 32 * i);
  }
  // one token row of the embedding-output gradient: requested by `load_row`, completed (LayerNorm-1 backward, dropout mask of
  // the forward's embedding dropout, vit.py:158) by `finish_row` -- warp-collective
  struct Row { float4 a[NV4]; uint2 dy[L1 ? NV4 : 1]; float4 x[L1 ? NV4 : 1]; float mu, rs; };
  auto load_row = [&](size_t r, Row& R) {
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      R.a[i] = __ldg(reinterpret_cast<const float4*>(dx + r * D) + lane + 32 * i);
      if constexpr (L1) {
        R.dy[i] = __ldg(reinterpret_cast<const uint2*>(l1.dy + r * D) + lane + 32 * i);
        R.x[i] = __ldg(reinterpret_cast<const float4*>(l1.x + r * D) + lane + 32 * i);
      }
    }
    if constexpr (L1) { R.mu = __ldg(l1.mean + r); R.rs = __ldg(l1.rstd + r); }
  };
  auto finish_row = [&](size_t r, Row& R) {
    if constexpr (L1) {
      float4 gg[NV4], xh[NV4];
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int i = 0; i < NV4; ++i) {
        const float2 lo = unpack_bf16x2(R.dy[i].x), hi = unpack_bf16x2(R.dy[i].y);
        const float4 d = make_float4(lo.x, lo.y, hi.x, hi.y), xv = R.x[i];
        xh[i] = make_float4((xv.x - R.mu) * R.rs, (xv.y - R.mu) * R.rs, (xv.z - R.mu) * R.rs, (xv.w - R.mu) * R.rs);
        gg[i] = make_float4(d.x * gm1[i].x, d.y * gm1[i].y, d.z * gm1[i].z, d.w * gm1[i].w);
        s1 += (gg[i].x + gg[i].y) + (gg[i].z + gg[i].w);
        s2 += (gg[i].x * xh[i].x + gg[i].y * xh[i].y) + (gg[i].z * xh[i].z + gg[i].w * xh[i].w);
        ag1[i].x += d.x * xh[i].x; ag1[i].y += d.y * xh[i].y; ag1[i].z += d.z * xh[i].z; ag1[i].w += d.w * xh[i].w;
        ab1[i].x += d.x; ab1[i].y += d.y; ab1[i].z += d.z; ab1[i].w += d.w;
      }
      s1 = warp_sum(s1) / (float)D;
      s2 = warp_sum(s2) / (float)D;
#pragma unroll
      for (int i = 0; i < NV4; ++i) {
        R.a[i].x += R.rs * (gg[i].x - s1 - xh[i].x * s2); R.a[i].y += R.rs * (gg[i].y - s1 - xh[i].y * s2);
        R.a[i].z += R.rs * (gg[i].z - s1 - xh[i].z * s2); R.a[i].w += R.rs * (gg[i].w - s1 - xh[i].w * s2);
      }
    }
    if (drop.thresh != 0) {
      const uint32_t rk = drop_row_key(drop, (uint32_t)r);
#pragma unroll
      for (int i = 0; i < NV4; ++i) {
        float dm[4];
        drop4(drop, rk, (uint32_t)(4 * (lane + 32 * i)), dm);
        R.a[i].x *= dm[0]; R.a[i].y *= dm[1]; R.a[i].z *= dm[2]; R.a[i].w *= dm[3];
      }
    }
  };
  // the warps' partial sums `acc` -> out0 (and out1): one atomicAdd per column and call
  auto flush = [&](const float4* acc, float* out0, float* out1) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV4; ++i) red[w][lane + 32 * i] = acc[i];
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += 256) {
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) t += reinterpret_cast<const float*>(red[k])[c];
      atomicAdd(out0 + c, t);
      if (out1 != nullptr) atomicAdd(out1 + c, t);
    }
  };
  struct Raw { Row a, u; float4 x[LN ? NV4 : 1]; float mu, rs; };
  const int items = S1 * nchunk;
#pragma unroll 1
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int j = item % S1, ch = item / S1;                 // neighbouring CTAs read neighbouring positions of the same samples
    const int b0 = ch * bchunk, b1 = min(B, b0 + bchunk);
    float4 acc[NV4];
    if (j == 0) {
      // token rows: mode 0 -> set 1 (exploitative pass; exploration_token, bug kept), mode 1 -> set 0, mode 2 -> both sets
      const int nset = mode == 2 ? 2 : 1;
      for (int q = 0; q < nset; ++q) {
        const int qs = mode == 0 ? 1 : q;
        zero(acc);
        for (int b = b0 + w; b < b1; b += 8) {
          const size_t r = row(qs, b, 0);
          Row R;
          load_row(r, R);
          finish_row(r, R);
#pragma unroll
          for (int i = 0; i < NV4; ++i) { acc[i].x += R.a[i].x; acc[i].y += R.a[i].y; acc[i].z += R.a[i].z; acc[i].w += R.a[i].w; }
        }
        flush(acc, dpos, q == 0 ? dtokA : dtokB);
      }
      continue;
    }
    zero(acc);
    auto load = [&](int b, Raw& R) {
      load_row(row(0, b, j), R.a);
      if (mode != 1) load_row(row(1, b, j), R.u);
      if constexpr (LN) {
        const size_t wid = (size_t)b * np + (j - 1);
#pragma unroll
        for (int i = 0; i < NV4; ++i) R.x[i] = __ldg(reinterpret_cast<const float4*>(ln.e0 + wid * D) + lane + 32 * i);
        R.mu = __ldg(ln.mean + wid); R.rs = __ldg(ln.rstd + wid);
      }
    };
    auto process = [&](int b, Raw& R) {
      const size_t wid = (size_t)b * np + (j - 1);
      finish_row(row(0, b, j), R.a);
      if (mode != 1) finish_row(row(1, b, j), R.u);
      float4 v[NV4];
#pragma unroll
      for (int i = 0; i < NV4; ++i) {
        const int c = lane + 32 * i;
        v[i] = R.a.a[i];
        float4 pz = v[i];                                  // what this position's dpos receives
        if (mode != 1) {
          const float4 u = R.u.a[i];
          v[i].x += u.x; v[i].y += u.y; v[i].z += u.z; v[i].w += u.w;
          pz = mode == 0 ? u : v[i];                       // mode 0: only the exploitative pass adds the positional embedding
        }
        acc[i].x += pz.x; acc[i].y += pz.y; acc[i].z += pz.z; acc[i].w += pz.w;
        if constexpr (!LN) {
          if (g != nullptr) *(reinterpret_cast<float4*>(g + wid * D) + c) = v[i];
          if (g_bf16 != nullptr)
            *(reinterpret_cast<uint2*>(g_bf16 + wid * D) + c) = make_uint2(pack_bf16x2(v[i].x, v[i].y), pack_bf16x2(v[i].z, v[i].w));
        }
      }
      if constexpr (LN) {
        // dx = rstd * (g*gamma - mean(g*gamma) - xhat * mean(g*gamma*xhat)), the same evaluation order as layernorm_bwd_kernel
        const float mu = R.mu, rs = R.rs;
        float4 gg[NV4], xh[NV4];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < NV4; ++i) {
          const float4 d = v[i], xv = R.x[i];
          xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
          gg[i] = make_float4(d.x * gm[i].x, d.y * gm[i].y, d.z * gm[i].z, d.w * gm[i].w);
          s1 += (gg[i].x + gg[i].y) + (gg[i].z + gg[i].w);
          s2 += (gg[i].x * xh[i].x + gg[i].y * xh[i].y) + (gg[i].z * xh[i].z + gg[i].w * xh[i].w);
          ag[i].x += d.x * xh[i].x; ag[i].y += d.y * xh[i].y; ag[i].z += d.z * xh[i].z; ag[i].w += d.w * xh[i].w;
          ab[i].x += d.x; ab[i].y += d.y; ab[i].z += d.z; ab[i].w += d.w;
        }
        s1 = warp_sum(s1) / (float)D;
        s2 = warp_sum(s2) / (float)D;
#pragma unroll
        for (int i = 0; i < NV4; ++i) {
          const float4 o = make_float4(rs * (gg[i].x - s1 - xh[i].x * s2), rs * (gg[i].y - s1 - xh[i].y * s2),
                                       rs * (gg[i].z - s1 - xh[i].z * s2), rs * (gg[i].w - s1 - xh[i].w * s2));
          ax[i].x += o.x; ax[i].y += o.y; ax[i].z += o.z; ax[i].w += o.w;
          *(reinterpret_cast<uint2*>(g_bf16 + wid * D) + lane + 32 * i) = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
        }
      }
    };
    if constexpr (L1) {
      // one sample per warp in flight: a sample is 6 KB of operands here (two token rows with their LayerNorm-1 operands)
      for (int b = b0 + w; b < b1; b += 8) {
        Raw A;
        load(b, A);
        process(b, A);
      }
    } else {
      // two samples per warp in flight: the second one's rows are requested before the first one is reduced
      for (int b = b0 + w; b < b1; b += 16) {
        Raw A, Bn;
        load(b, A);
        const bool two = b + 8 < b1;
        if (two) load(b + 8, Bn);
        process(b, A);
        if (two) process(b + 8, Bn);
      }
    }
    flush(acc, dpos + (size_t)j * D, nullptr);
  }
  if constexpr (LN) {
    flush(ag, ln.dgamma, nullptr);
    flush(ab, ln.dbeta, nullptr);
    if (ln.dbias != nullptr) flush(ax, ln.dbias, nullptr);
  }
  if constexpr (L1) {
    flush(ag1, l1.dgamma, nullptr);
    flush(ab1, l1.dbeta, nullptr);
  }
}


}  // namespace eavit

using namespace eavit;

extern "C" {

static int patchify_common(int mode, const void* img, int img_dtype, const long long* sample_idx, int B, int C, int HW, int P,
                           int order_cpp, const float* gamma, const float* beta, float eps, void* out, float* mean,
                           float* rstd, const float* dpln, float* dgamma, float* dbeta, cudaStream_t st) {
  EAVIT_CHECK_ARG(img && B > 0 && C > 0 && P > 0 && HW % P == 0);
  const int PD = C * P * P;
  EAVIT_CHECK_ARG(PD <= 32 * PATCH_MAXV);
  const int npr = HW / P;
  size_t smem = (size_t)C * P * HW * sizeof(float) + (size_t)PD * sizeof(int) + (mode == 1 ? (size_t)4 * 2 * PD * sizeof(float) : 0);
  EAVIT_CHECK_ARG(smem <= 48 * 1024);
  int nblk = B * npr;
  if (nblk > 16 * kNumSMs) nblk = 16 * kNumSMs;
  dim3 grid(nblk);
#define EAVIT_PATCHIFY(T, M, V) patchify_kernel<T, M, V><<<grid, 128, smem, st>>>((const T*)img, sample_idx, B, C, HW, P, order_cpp, gamma, beta, eps, (__nv_bfloat16*)out, mean, rstd, dpln, dgamma, dbeta)
#define EAVIT_PATCHIFY_V(T, M) do { if (PD <= 32 * 5) EAVIT_PATCHIFY(T, M, 5); else EAVIT_PATCHIFY(T, M, PATCH_MAXV); } while (0)
  if (img_dtype == EAVIT_U8) {
    if (mode == 0) EAVIT_PATCHIFY_V(uint8_t, 0); else EAVIT_PATCHIFY_V(uint8_t, 1);
  } else if (img_dtype == EAVIT_F32) {
    if (mode == 0) EAVIT_PATCHIFY_V(float, 0); else EAVIT_PATCHIFY_V(float, 1);
  } else { set_error("patchify: bad img dtype %d", img_dtype); return EAVIT_EINVAL; }
#undef EAVIT_PATCHIFY_V
#undef EAVIT_PATCHIFY
  EAVIT_LAUNCH_OK();
  return EAVIT_OK;
}

int eavit_patchify(const void* img, int img_dtype, const long long* sample_idx, int B, int C, int HW, int P, int order_cpp,
                   const float* gamma, const float* beta, float eps, void* out_bf16, float* mean, float* rstd,
                   void* stream) {
  EAVIT_CHECK_ARG(out_bf16 != nullptr);
  EAVIT_CHECK_ARG((gamma == nullptr) == (beta == nullptr));
  return patchify_common(0, img, img_dtype, sample_idx, B, C, HW, P, order_cpp, gamma, beta, eps, out_bf16, mean, rstd,
                         nullptr, nullptr, nullptr, (cudaStream_t)stream);
}

int eavit_patchify_ln_bwd(const void* img, int img_dtype, const long long* sample_idx, int B, int C, int HW, int P,
                          int order_cpp, const float* gamma, const float* mean, const float* rstd, const float* dpln,
                          float* dgamma, float* dbeta, void* stream) {
  EAVIT_CHECK_ARG(gamma && mean && rstd && dpln && dgamma && dbeta);
  return patchify_common(1, img, img_dtype, sample_idx, B, C, HW, P, order_cpp, gamma, gamma, 0.f, nullptr,
                         const_cast<float*>(mean), const_cast<float*>(rstd), dpln, dgamma, dbeta, (cudaStream_t)stream);
}

int eavit_embed_assemble(const float* e, const float* pos, const float* tokA, const float* tokB, int mode, int B, int np,
                         int D, float* x, void* stream) {
  EAVIT_CHECK_ARG(e && pos && tokA && x && B > 0 && np > 0 && D % 4 == 0 && mode >= 0 && mode <= 2);
  EAVIT_CHECK_ARG(mode != 2 || tokB != nullptr);
  embed_assemble_kernel<<<cdiv((long long)B * (np + 1), 8), 256, 0, (cudaStream_t)stream>>>(e, pos, tokA, tokB, mode, B, np, D, x);
  EAVIT_LAUNCH_OK();
  return EAVIT_OK;
}

// see include/eavit_b200.h: eavit_patch_ln_fold_bwd.  One CTA per input column k, thread j per output row (N <= 1024).
__global__ void __launch_bounds__(1024) patch_ln_fold_bwd_kernel(const float* __restrict__ G, const float* __restrict__ sv,
                                                                 const float* __restrict__ W, const float* __restrict__ g1,
                                                                 const float* __restrict__ b1, float* __restrict__ dW,
                                                                 float* __restrict__ dbias, float* __restrict__ dg1,
                                                                 float* __restrict__ db1, int N, int K) {
  __shared__ float ra[32], rb[32];
  const int k = blockIdx.x, j = threadIdx.x;
  float a = 0.f, b = 0.f;
  if (j < N) {
    const float gjk = G[(size_t)j * K + k], sj = sv[j], w = W[(size_t)j * K + k];
    dW[(size_t)j * K + k] += fmaf(g1[k], gjk, b1[k] * sj);
    if (k == 0) dbias[j] += sj;
    a = w * gjk; b = w * sj;
  }
  a = eavit::warp_sum(a); b = eavit::warp_sum(b);
  if ((j & 31) == 0) { ra[j >> 5] = a; rb[j >> 5] = b; }
  __syncthreads();
  if (j < 32) {
    const int nw = (blockDim.x + 31) >> 5;
    a = j < nw ? ra[j] : 0.f; b = j < nw ? rb[j] : 0.f;
    a = eavit::warp_sum(a); b = eavit::warp_sum(b);
    if (j == 0) { dg1[k] += a; db1[k] += b; }
  }
}

// work items of the one-pass embedding backward: (position, chunk of samples); chunks of 64 samples (8 per warp) unless that
// leaves fewer items than a few per co-resident CTA; a persistent grid of two CTAs per SM walks them
struct EabGrid { int bchunk, nchunk, grid; };
static EabGrid eab_grid(int B, int np) {
  EabGrid e;
  e.bchunk = 64;
  while (e.bchunk > 16 && (long long)(np + 1) * cdiv(B, e.bchunk) < 8LL * kNumSMs) e.bchunk >>= 1;
  e.nchunk = cdiv(B, e.bchunk);
  const long long items = (long long)(np + 1) * e.nchunk;
  e.grid = (int)(items < 2LL * kNumSMs ? items : 2LL * kNumSMs);
  return e;
}

int eavit_embed_assemble_bwd(const float* dx, int mode, int B, int np, int D, float* g, void* g_bf16, float* dpos,
                             float* dtokA, float* dtokB, void* stream) {
  EAVIT_CHECK_ARG(dx && (g || g_bf16) && dpos && dtokA && B > 0 && np > 0 && D % 4 == 0 && mode >= 0 && mode <= 2);
  EAVIT_CHECK_ARG(mode != 2 || dtokB != nullptr);
  cudaStream_t st = (cudaStream_t)stream;
  static const bool fused = getenv("EAVIT_NO_EMBED_BWD_FUSED") == nullptr;
  if (fused && D % 128 == 0 && D <= 1024) {
    // enough CTAs for a few waves, at least 8 samples per warp-pass so that the partial sums amortise their flush
    const EabGrid eg = eab_grid(B, np);
#define EAVIT_EAB(NV) embed_assemble_bwd_fused_kernel<NV, false, false><<<eg.grid, 256, 0, st>>>(dx, mode, B, np, g, (__nv_bfloat16*)g_bf16, dpos, dtokA, dtokB, eg.bchunk, eg.nchunk, EmbedLnBwd{}, EmbedLn1Bwd{}, make_drop(0.f, 0))
    switch (D / 128) {
      case 1: EAVIT_EAB(1); break; case 2: EAVIT_EAB(2); break; case 3: EAVIT_EAB(3); break; case 4: EAVIT_EAB(4); break;
      case 5: EAVIT_EAB(5); break; case 6: EAVIT_EAB(6); break; case 7: EAVIT_EAB(7); break; default: EAVIT_EAB(8); break;
    }
#undef EAVIT_EAB
    EAVIT_LAUNCH_OK();
    return EAVIT_OK;
  }
  embed_assemble_bwd_rows_kernel<<<cdiv((long long)B * np, 8), 256, 0, st>>>(dx, mode, B, np, D, g, (__nv_bfloat16*)g_bf16);
  EAVIT_LAUNCH_OK();
  dim3 grid(np + 1, cdiv(D, 128));
  embed_assemble_bwd_pos_kernel<<<grid, 128, 0, st>>>(dx, mode, B, np, D, dpos, dtokA, dtokB);
  EAVIT_LAUNCH_OK();
  return EAVIT_OK;
}

int eavit_patch_ln_fold_bwd(const float* G, const float* s, const float* W, const float* g1, const float* b1, float* dW,
                            float* dbias, float* dg1, float* db1, int N, int K, void* stream) {
  EAVIT_CHECK_ARG(G && s && W && g1 && b1 && dW && dbias && dg1 && db1 && N > 0 && N <= 1024 && K > 0);
  patch_ln_fold_bwd_kernel<<<K, ((N + 31) / 32) * 32, 0, (cudaStream_t)stream>>>(G, s, W, g1, b1, dW, dbias, dg1, db1, N, K);
  EAVIT_LAUNCH_OK();
  return EAVIT_OK;
}

int eavit_embed_assemble_ln_bwd(const float* dx, int mode, int B, int np, int D, const float* e0, const float* mean,
                                const float* rstd, const float* gamma, void* de_bf16, float* dgamma, float* dbeta, float* dbias,
                                float* dpos, float* dtokA, float* dtokB, float drop_p, unsigned long long drop_seed,
                                const void* l1_dy_bf16, const float* l1_x, const float* l1_mean, const float* l1_rstd,
                                const float* l1_gamma, float* l1_dgamma, float* l1_dbeta, void* stream) {
  EAVIT_CHECK_ARG(dx && e0 && mean && rstd && gamma && de_bf16 && dgamma && dbeta && dpos && dtokA && B > 0 && np > 0);
  EAVIT_CHECK_ARG(l1_dy_bf16 == nullptr || (l1_x && l1_mean && l1_rstd && l1_gamma && l1_dgamma && l1_dbeta));
  EAVIT_CHECK_ARG(mode >= 0 && mode <= 2 && (mode != 2 || dtokB != nullptr));
  EAVIT_CHECK_ARG(D == 256);                         // the row a warp holds (2 x float4 per lane); other widths: the two separate calls
  const EabGrid eg = eab_grid(B, np);
  const EmbedLnBwd ln{e0, mean, rstd, gamma, dgamma, dbeta, dbias};
  const EmbedLn1Bwd l1{(const __nv_bfloat16*)l1_dy_bf16, l1_x, l1_mean, l1_rstd, l1_gamma, l1_dgamma, l1_dbeta};
  if (l1_dy_bf16 != nullptr)
    embed_assemble_bwd_fused_kernel<2, true, true><<<eg.grid, 256, 0, (cudaStream_t)stream>>>(dx, mode, B, np, nullptr, (__nv_bfloat16*)de_bf16,
                                                                                             dpos, dtokA, dtokB, eg.bchunk, eg.nchunk, ln, l1,
                                                                                             make_drop(drop_p, drop_seed));
  else
    embed_assemble_bwd_fused_kernel<2, true, false><<<eg.grid, 256, 0, (cudaStream_t)stream>>>(dx, mode, B, np, nullptr, (__nv_bfloat16*)de_bf16,
                                                                                              dpos, dtokA, dtokB, eg.bchunk, eg.nchunk, ln, l1,
                                                                                              make_drop(drop_p, drop_seed));
  EAVIT_LAUNCH_OK();
  return EAVIT_OK;
}

}  // extern "C"
