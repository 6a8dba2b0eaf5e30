// LayerNorm forward / backward over rows of a [T, D] fp32 residual stream (one warp per row, float4 loads),
// bf16 output feeding the tcgen05 GEMMs; backward fuses the residual-path add and the dgamma/dbeta
// reductions.  Also: bf16 column sums (bias gradients) and row gather/scatter for the pooled token.
//
// Reference: nn.LayerNorm at vit.py:28 (FeedForward), :47 (Attention), :78 (Transformer.norm), :111/:113
// (patch embedding); HF layernorm_before/after/final (vit_hg.py via transformers).
#include "common.cuh"

namespace eavit {

constexpr int LN_MAXV = 8;   // D <= 32 lanes * 4 * 8 = 1024

// y = (x - mean) * rstd * gamma + beta ; two-pass statistics in fp32 like ATen's CPU/CUDA kernels.
template <typename OutT, int VPL>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const float* __restrict__ x, long long ldx,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, OutT* __restrict__ y,
                                                            long long ldy, float* __restrict__ mean_out,
                                                            float* __restrict__ rstd_out, int T, int D, float eps) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= T) return;
  const float* xr = x + (size_t)row * ldx;
  const int nv = D / 4;
  float4 v[VPL];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = lane + 32 * i;
    if (c < nv) {
      v[i] = __ldg(reinterpret_cast<const float4*>(xr) + c);
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
  const float mean = warp_sum(s) / (float)D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = lane + 32 * i;
    if (c < nv) {
      const float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (cc * cc + d * d);
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
  if (lane == 0 && mean_out != nullptr) { mean_out[row] = mean; rstd_out[row] = rstd; }
  OutT* yr = y + (size_t)row * ldy;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = lane + 32 * i;
    if (c < nv) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c);
      const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + c);
      const float o0 = (v[i].x - mean) * rstd * g.x + b.x, o1 = (v[i].y - mean) * rstd * g.y + b.y;
      const float o2 = (v[i].z - mean) * rstd * g.z + b.z, o3 = (v[i].w - mean) * rstd * g.w + b.w;
      if constexpr (sizeof(OutT) == 4) {
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(yr) + 4 * c) = make_float4(o0, o1, o2, o3);
      } else {
        *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(yr) + 4 * c) =
            make_uint2(pack_bf16x2(o0, o1), pack_bf16x2(o2, o3));
      }
    }
  }
}

// dx = dres + rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma
// dgamma += sum_rows dy * xhat ; dbeta += sum_rows dy      (block partials -> one atomicAdd per column per CTA)
template <typename DyT, int VPL>
__global__ void __launch_bounds__(256, (VPL <= 2 ? 2 : 1)) layernorm_bwd_kernel(const DyT* __restrict__ dy, long long lddy,
                                                            const float* __restrict__ x, long long ldx,
                                                            const float* __restrict__ mean,
                                                            const float* __restrict__ rstd,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ dres, long long lddres,
                                                            float* __restrict__ dx, long long lddx,
                                                            __nv_bfloat16* __restrict__ dx_bf16, long long lddxb,
                                                            float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                            float* __restrict__ dxsum, const DropCfg drop, int T, int D) {
  extern __shared__ float s_part[];   // [8 warps][3][D]
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nv = D / 4;
  float4 ag[VPL], ab[VPL], ax[VPL];   // dgamma, dbeta, column sums of the output dx (bias gradient of the next GEMM)
#pragma unroll
  for (int i = 0; i < VPL; ++i) { ag[i] = make_float4(0.f, 0.f, 0.f, 0.f); ab[i] = ag[i]; ax[i] = ag[i]; }
  // persistent grid-stride over rows (the dgamma/dbeta atomics are paid once per CTA), with the NEXT row's operands
  // loaded into a second register set before the current row's reductions: a warp keeps two rows (5 KB at D = 256)
  // in flight instead of alternating between a load phase and a shuffle/compute phase
  float4 gm[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = lane + 32 * i;
    gm[i] = c < nv ? __ldg(reinterpret_cast<const float4*>(gamma) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  constexpr bool kPrefetch = VPL <= 2;       // wider rows would not fit two register sets
  struct Raw { float4 xv[VPL], r[VPL]; float mu, rs; float4 d32[sizeof(DyT) == 4 ? VPL : 1]; uint2 d16[sizeof(DyT) == 4 ? 1 : VPL]; };
  auto load = [&](int row, Raw& R) {
    R.mu = mean[row]; R.rs = rstd[row];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = lane + 32 * i;
      if (c < nv) {
        if constexpr (sizeof(DyT) == 4) {
          R.d32[i] = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dy) + (size_t)row * lddy) + c);
        } else {
          R.d16[i] = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(dy) + (size_t)row * lddy) + c);
        }
        R.xv[i] = __ldg(reinterpret_cast<const float4*>(x + (size_t)row * ldx) + c);
        if (dres != nullptr) R.r[i] = __ldg(reinterpret_cast<const float4*>(dres + (size_t)row * lddres) + c);
      }
    }
  };
  auto process = [&](int row, const Raw& R) {
    const float mu = R.mu, rs = R.rs;
    float4 g[VPL], xh[VPL];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = lane + 32 * i;
      if (c < nv) {
        float4 d;
        if constexpr (sizeof(DyT) == 4) {
          d = R.d32[i];
        } else {
          const float2 lo = unpack_bf16x2(R.d16[i].x), hi = unpack_bf16x2(R.d16[i].y);
          d = make_float4(lo.x, lo.y, hi.x, hi.y);
        }
        const float4 xv = R.xv[i];
        xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
        g[i] = make_float4(d.x * gm[i].x, d.y * gm[i].y, d.z * gm[i].z, d.w * gm[i].w);
        s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
        s2 += (g[i].x * xh[i].x + g[i].y * xh[i].y) + (g[i].z * xh[i].z + g[i].w * xh[i].w);
        ag[i].x += d.x * xh[i].x; ag[i].y += d.y * xh[i].y; ag[i].z += d.z * xh[i].z; ag[i].w += d.w * xh[i].w;
        ab[i].x += d.x; ab[i].y += d.y; ab[i].z += d.z; ab[i].w += d.w;
      }
    }
    s1 = warp_sum(s1) / (float)D;
    s2 = warp_sum(s2) / (float)D;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = lane + 32 * i;
      if (c < nv) {
        float4 o = make_float4(rs * (g[i].x - s1 - xh[i].x * s2), rs * (g[i].y - s1 - xh[i].y * s2),
                               rs * (g[i].z - s1 - xh[i].z * s2), rs * (g[i].w - s1 - xh[i].w * s2));
        if (dres != nullptr) { o.x += R.r[i].x; o.y += R.r[i].y; o.z += R.r[i].z; o.w += R.r[i].w; }
        if (dx != nullptr) *(reinterpret_cast<float4*>(dx + (size_t)row * lddx) + c) = o;
        // The bf16 copy (and its column sums) is the gradient w.r.t. the OUTPUT of the Linear that fed this residual
        // stream: with dropout between that Linear and the residual add (vit.py:33,56) it carries the forward mask.
        if (drop.thresh != 0) {
          float dm[4];
          drop4(drop, drop_row_key(drop, (uint32_t)row), (uint32_t)(4 * c), dm);
          o.x *= dm[0]; o.y *= dm[1]; o.z *= dm[2]; o.w *= dm[3];
        }
        ax[i].x += o.x; ax[i].y += o.y; ax[i].z += o.z; ax[i].w += o.w;
        if (dx_bf16 != nullptr)
          *(reinterpret_cast<uint2*>(dx_bf16 + (size_t)row * lddxb) + c) = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
      }
    }
  };
  {
    const int stride = gridDim.x * 8;
    int row = blockIdx.x * 8 + w;
    if constexpr (kPrefetch) {
      Raw A, B;
      if (row < T) load(row, A);
      while (row < T) {
        int nrow = row + stride;
        if (nrow < T) load(nrow, B);
        process(row, A);
        row = nrow;
        if (row >= T) break;
        nrow = row + stride;
        if (nrow < T) load(nrow, A);
        process(row, B);
        row = nrow;
      }
    } else {
      for (; row < T; row += stride) {
        Raw A;
        load(row, A);
        process(row, A);
      }
    }
  }
  if (dgamma == nullptr && dxsum == nullptr) return;
  float* pg = s_part + (size_t)w * 3 * D;
  float* pb = pg + D;
  float* px = pb + D;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = lane + 32 * i;
    if (c < nv) {
      *(reinterpret_cast<float4*>(pg) + c) = ag[i];
      *(reinterpret_cast<float4*>(pb) + c) = ab[i];
      *(reinterpret_cast<float4*>(px) + c) = ax[i];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float sg = 0.f, sb = 0.f, sx = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      sg += s_part[(size_t)k * 3 * D + c]; sb += s_part[(size_t)k * 3 * D + D + c]; sx += s_part[(size_t)k * 3 * D + 2 * D + c];
    }
    if (dgamma != nullptr) { atomicAdd(dgamma + c, sg); atomicAdd(dbeta + c, sb); }
    if (dxsum != nullptr) atomicAdd(dxsum + c, sx);
  }
}

// out[c] += sum_rows x[r, c]   (bias gradients); x bf16 or fp32 [T, N]
template <typename XT>
__global__ void __launch_bounds__(256) colsum_kernel(const XT* __restrict__ x, long long ldx, float* __restrict__ out,
                                                     int T, int N, int rows_per_block) {
  // thread (tx, ty): tx covers 2 columns, ty strides rows
  const int cpair = blockIdx.x * 32 + (threadIdx.x & 31);
  const int ty = threadIdx.x >> 5;
  const int c = cpair * 2;
  __shared__ float2 s[8][32];
  float a0 = 0.f, a1 = 0.f;
  if (c < N) {
    const int r0 = blockIdx.y * rows_per_block, r1 = min(T, r0 + rows_per_block);
    for (int r = r0 + ty; r < r1; r += 8) {
      if constexpr (sizeof(XT) == 2) {
        const float2 v = unpack_bf16x2(__ldg(reinterpret_cast<const uint32_t*>(reinterpret_cast<const __nv_bfloat16*>(x) + (size_t)r * ldx + c)));
        a0 += v.x; a1 += v.y;
      } else {
        const float2 v = __ldg(reinterpret_cast<const float2*>(reinterpret_cast<const float*>(x) + (size_t)r * ldx + c));
        a0 += v.x; a1 += v.y;
      }
    }
  }
  s[ty][threadIdx.x & 31] = make_float2(a0, a1);
  __syncthreads();
  if (ty == 0 && c < N) {
#pragma unroll
    for (int k = 1; k < 8; ++k) { a0 += s[k][threadIdx.x].x; a1 += s[k][threadIdx.x].y; }
    atomicAdd(out + c, a0);
    atomicAdd(out + c + 1, a1);
  }
}

// dst[i, :] = src[idx(i), :] with idx(i) = rows[i]  (gather of the pooled token rows; fp32)
__global__ void gather_rows_kernel(const float* __restrict__ src, long long lds, const int* __restrict__ rows,
                                   float* __restrict__ dst, long long ldd, int n, int D) {
  const int i = blockIdx.x;
  const float* s = src + (size_t)rows[i] * lds;
  for (int c = threadIdx.x; c < D / 4; c += blockDim.x)
    *(reinterpret_cast<float4*>(dst + (size_t)i * ldd) + c) = __ldg(reinterpret_cast<const float4*>(s) + c);
}
// dst[rows[i], :] = src[i, :]  (scatter; rows are unique)  + optional bf16 copy
__global__ void scatter_rows_kernel(const float* __restrict__ src, long long lds, const int* __restrict__ rows,
                                    float* __restrict__ dst, long long ldd, __nv_bfloat16* __restrict__ dst_bf16,
                                    long long lddb, int n, int D) {
  const int i = blockIdx.x;
  const size_t r = (size_t)rows[i];
  for (int c = threadIdx.x; c < D / 4; c += blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(src + (size_t)i * lds) + c);
    if (dst != nullptr) *(reinterpret_cast<float4*>(dst + r * ldd) + c) = v;
    if (dst_bf16 != nullptr)
      *(reinterpret_cast<uint2*>(dst_bf16 + r * lddb) + c) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  }
}

// dst[rows[i], :] += src[i, :] for a few rows of a dense gradient (the pooled tokens' residual gradient of the pruned last
// layer), keeping the two by-products of the LayerNorm backward that produced dst consistent: the bf16 copy of the rows
// (under the dropout mask of the linear layer below) and that layer's bias gradient (+= column sums of the masked delta).
// Replaces zero-filling a [T, D] fp32 buffer, scattering into it and streaming it through the LayerNorm backward as `dres`.
__global__ void __launch_bounds__(64) scatter_add_rows_kernel(const float* __restrict__ src, long long lds, const int* __restrict__ rows,
                                                              float* __restrict__ dst, long long ldd, __nv_bfloat16* __restrict__ dst_bf16,
                                                              long long lddb, float* __restrict__ colsum, int n, int D, int rows_per_block,
                                                              const DropCfg drop) {
  const int i0 = blockIdx.x * rows_per_block, i1 = min(n, i0 + rows_per_block);
  for (int c = threadIdx.x; c < D / 4; c += blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = i0; i < i1; ++i) {
      const size_t r = (size_t)rows[i];
      const float4 d = __ldg(reinterpret_cast<const float4*>(src + (size_t)i * lds) + c);
      float4 v = *(reinterpret_cast<float4*>(dst + r * ldd) + c);
      v.x += d.x; v.y += d.y; v.z += d.z; v.w += d.w;
      *(reinterpret_cast<float4*>(dst + r * ldd) + c) = v;
      float dm[4] = {1.f, 1.f, 1.f, 1.f};
      if (drop.thresh) drop4(drop, drop_row_key(drop, (uint32_t)r), (uint32_t)(4 * c), dm);
      if (dst_bf16 != nullptr)
        *(reinterpret_cast<uint2*>(dst_bf16 + r * lddb) + c) = make_uint2(pack_bf16x2(v.x * dm[0], v.y * dm[1]), pack_bf16x2(v.z * dm[2], v.w * dm[3]));
      acc.x += d.x * dm[0]; acc.y += d.y * dm[1]; acc.z += d.z * dm[2]; acc.w += d.w * dm[3];
    }
    if (colsum != nullptr) {
      atomicAdd(colsum + 4 * c, acc.x); atomicAdd(colsum + 4 * c + 1, acc.y);
      atomicAdd(colsum + 4 * c + 2, acc.z); atomicAdd(colsum + 4 * c + 3, acc.w);
    }
  }
}

// dst[i, c] = src[i, c] * dropmask(rows ? rows[i] : row0 + i, c)   (embedding dropout, vit.py:158; top-gradient rows)
__global__ void __launch_bounds__(256) dropout_apply_kernel(const float* __restrict__ src, long long lds, const int* __restrict__ rows,
                                                            int row0, float* __restrict__ dst, long long ldd, int n, int D,
                                                            const DropCfg drop) {
  const int nv = D / 4;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < (long long)n * nv; e += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(e / nv), c = (int)(e - (long long)i * nv);
    const uint32_t r = rows ? (uint32_t)rows[i] : (uint32_t)(row0 + i);
    float4 v = *(reinterpret_cast<const float4*>(src + (size_t)i * lds) + c);
    float dm[4];
    drop4(drop, drop_row_key(drop, r), (uint32_t)(4 * c), dm);
    v.x *= dm[0]; v.y *= dm[1]; v.z *= dm[2]; v.w *= dm[3];
    *(reinterpret_cast<float4*>(dst + (size_t)i * ldd) + c) = v;
  }
}
__global__ void __launch_bounds__(256) dropout_mask_kernel(float* __restrict__ out, long long ld, int row0, int n, int col0, int ncols,
                                                           const DropCfg drop) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < (long long)n * ncols; e += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(e / ncols), j = (int)(e - (long long)i * ncols);
    const uint32_t c = (uint32_t)(col0 + j);
    const uint32_t bits = drop_bits(drop_row_key(drop, (uint32_t)(row0 + i)), c >> 1);
    out[(size_t)i * ld + j] = (c & 1) ? drop_odd(drop, bits) : drop_even(drop, bits);
  }
}

// fp32 -> bf16 elementwise (weight shadows, activations)
__global__ void cast_f32_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n4) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(in) + i);
    *(reinterpret_cast<uint2*>(out) + i) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  }
}

}  // namespace eavit

using namespace eavit;

extern "C" {

int eavit_layernorm_fwd(const float* x, long long ldx, const float* gamma, const float* beta, void* y, int y_dtype,
                        long long ldy, float* mean, float* rstd, int T, int D, float eps, void* stream) {
  EAVIT_CHECK_ARG(T > 0 && D > 0 && D % 4 == 0 && D <= 128 * LN_MAXV && x && gamma && beta && y);
  EAVIT_CHECK_ARG(ldx % 4 == 0 && ldy % 4 == 0);
  EAVIT_CHECK_ARG((mean == nullptr) == (rstd == nullptr));
  cudaStream_t st = (cudaStream_t)stream;
  const int vpl = cdiv(D / 4, 32);            // float4 vectors per lane: registers (and occupancy) sized for the actual D
#define EAVIT_LN_FWD(OT, V) layernorm_fwd_kernel<OT, V><<<cdiv(T, 8), 256, 0, st>>>(x, ldx, gamma, beta, (OT*)y, ldy, mean, rstd, T, D, eps)
#define EAVIT_LN_FWD_V(OT) do { if (vpl <= 1) EAVIT_LN_FWD(OT, 1); else if (vpl <= 2) EAVIT_LN_FWD(OT, 2); else if (vpl <= 4) EAVIT_LN_FWD(OT, 4); else EAVIT_LN_FWD(OT, 8); } while (0)
  if (y_dtype == EAVIT_BF16) EAVIT_LN_FWD_V(__nv_bfloat16);
  else if (y_dtype == EAVIT_F32) EAVIT_LN_FWD_V(float);
  else { set_error("layernorm_fwd: bad y_dtype %d", y_dtype); return EAVIT_EINVAL; }
#undef EAVIT_LN_FWD_V
#undef EAVIT_LN_FWD
  EAVIT_LAUNCH_OK();
  return EAVIT_OK;
}

int eavit_layernorm_bwd(const void* dy, int dy_dtype, long long lddy, const float* x, long long ldx, const float* mean,
                        const float* rstd, const float* gamma, const float* dres, long long lddres, float* dx,
                        long long lddx, void* dx_bf16, long long lddxb, float* dgamma, float* dbeta, float* dxsum, float drop_p,
                        unsigned long long drop_seed, int T, int D, void* stream) {
  EAVIT_CHECK_ARG(T > 0 && D > 0 && D % 4 == 0 && D <= 128 * LN_MAXV && dy && x && mean && rstd && gamma);
  EAVIT_CHECK_ARG((dgamma == nullptr) == (dbeta == nullptr));
  EAVIT_CHECK_ARG(lddy % 4 == 0 && ldx % 4 == 0 && lddres % 4 == 0 && lddx % 4 == 0 && lddxb % 4 == 0);
  EAVIT_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f);
  const DropCfg drop = make_drop(drop_p, drop_seed);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = (size_t)8 * 3 * D * sizeof(float);
  const int vpl = cdiv(D / 4, 32);
  static bool attr_done = false;
  if (!attr_done) {
    EAVIT_CUDA(cudaFuncSetAttribute(layernorm_bwd_kernel<float, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 3 * 1024 * 4));
    EAVIT_CUDA(cudaFuncSetAttribute(layernorm_bwd_kernel<__nv_bfloat16, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 3 * 1024 * 4));
    attr_done = true;
  }
  // persistent grid = exactly the CTAs that are co-resident (a partial second wave would run at a fraction of the bandwidth)
#define EAVIT_LN_BWD(DT, V) do { int occ = 1; EAVIT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, layernorm_bwd_kernel<DT, V>, 256, smem)); \
    int grid = cdiv(T, 8); if (occ < 1) occ = 1; if (grid > occ * kNumSMs) grid = occ * kNumSMs; \
    layernorm_bwd_kernel<DT, V><<<grid, 256, smem, st>>>((const DT*)dy, lddy, x, ldx, mean, rstd, gamma, dres, lddres, dx, lddx, (__nv_bfloat16*)dx_bf16, lddxb, dgamma, dbeta, dxsum, drop, T, D); } while (0)
#define EAVIT_LN_BWD_V(DT) do { if (vpl <= 1) EAVIT_LN_BWD(DT, 1); else if (vpl <= 2) EAVIT_LN_BWD(DT, 2); else if (vpl <= 4) EAVIT_LN_BWD(DT, 4); else EAVIT_LN_BWD(DT, 8); } while (0)
  if (dy_dtype == EAVIT_F32) EAVIT_LN_BWD_V(float);
  else if (dy_dtype == EAVIT_BF16) EAVIT_LN_BWD_V(__nv_bfloat16);
  else { set_error("layernorm_bwd: bad dy_dtype %d", dy_dtype); return EAVIT_EINVAL; }
#undef EAVIT_LN_BWD_V
#undef EAVIT_LN_BWD
  EAVIT_LAUNCH_OK();
  return EAVIT_OK;
}

int eavit_colsum(const void* x, int x_dtype, long long ldx, float* out, int T, int N, void* stream) {
  EAVIT_CHECK_ARG(T > 0 && N > 0 && N % 2 == 0 && x && out && ldx % 2 == 0);
  cudaStream_t st = (cudaStream_t)stream;
  const int col_blocks = cdiv(N, 64);
  int row_blocks = cdiv(4 * kNumSMs, col_blocks);
  if (row_blocks > cdiv(T, 8)) row_blocks = cdiv(T, 8);
  const int rpb = cdiv(T, row_blocks);
  dim3 grid(col_blocks, cdiv(T, rpb));
  if (x_dtype == EAVIT_BF16) colsum_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, ldx, out, T, N, rpb);
  else if (x_dtype == EAVIT_F32) colsum_kernel<float><<<grid, 256, 0, st>>>((const float*)x, ldx, out, T, N, rpb);
  else { set_error("colsum: bad dtype %d", x_dtype); return EAVIT_EINVAL; }
  EAVIT_LAUNCH_OK();
  return EAVIT_OK;
}

int eavit_gather_rows(const float* src, long long lds, const int* rows, float* dst, long long ldd, int n, int D, void* stream) {
  EAVIT_CHECK_ARG(n > 0 && D % 4 == 0 && src && rows && dst && lds % 4 == 0 && ldd % 4 == 0);
  gather_rows_kernel<<<n, 64, 0, (cudaStream_t)stream>>>(src, lds, rows, dst, ldd, n, D);
  EAVIT_LAUNCH_OK();
  return EAVIT_OK;
}

int eavit_dropout_apply(const float* src, long long lds, const int* rows, int row0, float* dst, long long ldd, int n, int D,
                        float p, unsigned long long seed, void* stream) {
  EAVIT_CHECK_ARG(src && dst && n > 0 && D > 0 && D % 4 == 0 && lds % 4 == 0 && ldd % 4 == 0 && p >= 0.f && p < 1.f);
  long long blocks = ((long long)n * (D / 4) + 255) / 256;
  if (blocks > 16 * kNumSMs) blocks = 16 * kNumSMs;
  dropout_apply_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(src, lds, rows, row0, dst, ldd, n, D, make_drop(p, seed));
  EAVIT_LAUNCH_OK();
  return EAVIT_OK;
}

int eavit_dropout_mask(float* out, long long ld, int row0, int n, int col0, int ncols, float p, unsigned long long seed,
                       void* stream) {
  EAVIT_CHECK_ARG(out && n > 0 && ncols > 0 && p >= 0.f && p < 1.f);
  long long blocks = ((long long)n * ncols + 255) / 256;
  if (blocks > 16 * kNumSMs) blocks = 16 * kNumSMs;
  dropout_mask_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(out, ld, row0, n, col0, ncols, make_drop(p, seed));
  EAVIT_LAUNCH_OK();
  return EAVIT_OK;
}

int eavit_scatter_rows(const float* src, long long lds, const int* rows, float* dst, long long ldd, void* dst_bf16,
                       long long lddb, int n, int D, void* stream) {
  EAVIT_CHECK_ARG(n > 0 && D % 4 == 0 && src && rows && (dst || dst_bf16) && lds % 4 == 0 && ldd % 4 == 0 && lddb % 4 == 0);
  scatter_rows_kernel<<<n, 64, 0, (cudaStream_t)stream>>>(src, lds, rows, dst, ldd, (__nv_bfloat16*)dst_bf16, lddb, n, D);
  EAVIT_LAUNCH_OK();
  return EAVIT_OK;
}

int eavit_scatter_add_rows(const float* src, long long lds, const int* rows, float* dst, long long ldd, void* dst_bf16, long long lddb,
                           float* colsum, int n, int D, float drop_p, unsigned long long drop_seed, void* stream) {
  EAVIT_CHECK_ARG(n > 0 && D % 4 == 0 && src && rows && dst && lds % 4 == 0 && ldd % 4 == 0 && lddb % 4 == 0);
  EAVIT_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f);
  const int rpb = 32;
  scatter_add_rows_kernel<<<cdiv(n, rpb), 64, 0, (cudaStream_t)stream>>>(src, lds, rows, dst, ldd, (__nv_bfloat16*)dst_bf16, lddb, colsum, n, D,
                                                                        rpb, make_drop(drop_p, drop_seed));
  EAVIT_LAUNCH_OK();
  return EAVIT_OK;
}

int eavit_cast_f32_bf16(const float* in, void* out, long long n, void* stream) {
  EAVIT_CHECK_ARG(n > 0 && n % 4 == 0 && in && out);
  cast_f32_bf16_kernel<<<cdiv(n / 4, 256), 256, 0, (cudaStream_t)stream>>>(in, (__nv_bfloat16*)out, n / 4);
  EAVIT_LAUNCH_OK();
  return EAVIT_OK;
}

}  // extern "C"
