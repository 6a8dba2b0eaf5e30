// The lucidrains patch embedding as ONE kernel (north-star clause 1): to_patch_embedding (vit.py:109-114: Rearrange,
// LayerNorm(patch_dim), Linear(patch_dim, dim), LayerNorm(dim)) + token prepend + positional add (vit.py:141-158), for
// patch_dim <= 192, dim == 256:
//
//   producers (8 warps)   gather a patch (C x P runs of P contiguous pixels; uint8 frames are divided by 255 here) straight from
//                         global memory, LayerNorm(patch_dim) one patch per warp, and write the normalised bf16 patch (a) into the
//                         128-byte-swizzled A tile of the tensor-core GEMM [128 patches x 192] and (b) to global memory for the
//                         backward (the dW GEMM's operand) together with (mean, rstd)
//   warp 0                loads the whole weight [256, patch_dim] ONCE per CTA with TMA (3 boxes of 64 x 256, out-of-bounds
//                         columns zero-filled) and issues tcgen05.mma 128 x 256 x 16, 9 k-steps per tile, accumulator in TMEM
//   epilogue (4 warps)    a thread owns one patch = one TMEM lane: + bias, row statistics of LayerNorm(dim) without any
//                         cross-thread reduction, then per 32-column chunk a transpose through shared memory so that global
//                         stores are row-contiguous: e0 (the LayerNorm input, kept for the backward), the normalised token into
//                         the explorative sequence (no positional embedding -- the reference's token bug, vit.py:142) and
//                         token + pos_embedding[1 + j] into the exploitative / CLS sequence
// A tiles and accumulators are double-buffered: patches of tile i+1 are gathered while tile i is multiplied and normalised.
// Replaces four launches (patchify + LN, GEMM, LayerNorm, assemble) that moved ~680 MB per 512 samples with one that moves
// ~350 MB.  The token rows (token + pos_embedding[0]) are written by the producers before the first tile.
#include "common.cuh"
#include "tc05.cuh"

namespace eavit {

int make_tmap_bf16_2d(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_bytes, uint32_t box_outer);

namespace ef {

constexpr int D = 256;                       // model width == MMA N
constexpr int KPAD = 192;                    // patch_dim padded to 3 swizzle atoms of 64 bf16
constexpr int TILE_M = 128;
constexpr int PROD_WARPS = 8, EPI_WARPS = 4;
constexpr int THREADS = (1 + PROD_WARPS + EPI_WARPS) * 32;     // 416
constexpr int A_BYTES = 3 * TILE_M * 128;    // 3 k-blocks x [128 rows][128 B]
constexpr int W_BYTES = 3 * D * 128;         // 3 k-blocks x [256 rows][128 B]

constexpr int NSTAGE = 4;                    // ring of uint8 patch-row staging buffers (cp.async, prefetch distance 2)
constexpr int STAGE_BYTES = 2048;            // C * P rows of HW bytes (4 x 6 x 84 = 2016)

struct Smem {
  static constexpr int OFF_W = 0;
  static constexpr int OFF_A = OFF_W + W_BYTES;                    // 2 buffers
  static constexpr int OFF_EPI = OFF_A + 2 * A_BYTES;              // 4 warps x [32][32] fp32 transpose tiles
  static constexpr int OFF_STAGE = OFF_EPI + EPI_WARPS * 4096;     // uint8 [NSTAGE][C*P*HW] image rows of a patch row
  static constexpr int OFF_VEC = OFF_STAGE + NSTAGE * STAGE_BYTES; // float bias[256], g3[256], b3[256]
  static constexpr int OFF_LUT = OFF_VEC + 3 * D * 4;              // float [256]: b / 255 correctly rounded (np.float32(x) / 255.)
  static constexpr int OFF_BAR = OFF_LUT + 256 * 4;
  static constexpr int TOTAL = OFF_BAR + 128 + 1024;
};
static_assert(Smem::TOTAL <= 232448, "shared memory");

__device__ __forceinline__ uint32_t sw_off(int r, int c) {          // 16-byte chunk c of row r in a [rows][128 B] swizzled slab
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
}
__device__ __forceinline__ void prod_bar() { asm volatile("bar.sync 1, %0;" ::"n"(PROD_WARPS * 32) : "memory"); }

struct Params {
  const void* img; int img_u8; const long long* sample_idx;
  int B, C, HW, P, PD, np, npr, mode;          // np = patches per sample, npr = patches per image row
  const float *g1, *b1; float eps1;            // LayerNorm(patch_dim)
  const float* bias;                           // Linear bias [256]
  const float *g3, *b3; float eps3;            // LayerNorm(dim)
  const float *pos, *tok;                      // pos_embedding [np + 1, 256], exploration / cls token [256]
  __nv_bfloat16* pln; float *pmean, *prstd;    // saved for the backward
  float* e0; float *m3, *r3;
  float* x0;                                   // flat residual stream
  int rows, tiles;
  int pln_xhat;                                // pln receives the normalised patch WITHOUT the LayerNorm's affine (see the ABI comment)
};

template <typename ImgT>
__device__ __forceinline__ float pix(const ImgT* p) {
  if constexpr (sizeof(ImgT) == 1) return __fdiv_rn((float)(*p), 255.0f);
  else return *p;
}

template <typename ImgT>
__global__ void __launch_bounds__(THREADS, 1) embed_fused_kernel(const __grid_constant__ CUtensorMap tmW, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic on the __shared__ array keeps the address space: LDS / STS, not generic LD / ST
  float* s_bias = reinterpret_cast<float*>(smem + Smem::OFF_VEC);
  float* s_g3 = s_bias + D;
  float* s_b3 = s_g3 + D;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem::OFF_BAR);
  uint64_t* w_full = bars;            // [1]
  uint64_t* a_full = bars + 1;        // [2] producers -> MMA
  uint64_t* a_empty = bars + 3;       // [2] MMA -> producers
  uint64_t* acc_full = bars + 5;      // [2] MMA -> epilogue
  uint64_t* acc_empty = bars + 7;     // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // zero the A tiles once: the K padding (patch_dim .. 192) must be finite zeros for the MMA
  for (int i = threadIdx.x; i < 2 * A_BYTES / 16; i += THREADS) reinterpret_cast<uint4*>(smem + Smem::OFF_A)[i] = make_uint4(0, 0, 0, 0);
  for (int i = threadIdx.x; i < D; i += THREADS) { s_bias[i] = p.bias[i]; s_g3[i] = p.g3[i]; s_b3[i] = p.b3[i]; }
  float* s_lut = reinterpret_cast<float*>(smem + Smem::OFF_LUT);
  for (int i = threadIdx.x; i < 256; i += THREADS) s_lut[i] = __fdiv_rn((float)i, 255.0f);      // one IEEE division per value, not per pixel
  if (threadIdx.x == 0) {
    tc::prefetch_tmap(&tmW);
    tc::mbar_init(w_full, 1);
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&a_full[i], PROD_WARPS); tc::mbar_init(&a_empty[i], 1);
      tc::mbar_init(&acc_full[i], 1); tc::mbar_init(&acc_empty[i], EPI_WARPS);
    }
    tc::fence_barrier_init();
  }
  if (warp == 0) tc::tmem_alloc(tmem_slot, 512);
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================== weight load + MMA issue =====================================
    if (lane == 0) {
      tc::mbar_expect_tx(w_full, W_BYTES);
      for (int kb = 0; kb < 3; ++kb) tc::tma_load_2d(smem + Smem::OFF_W + kb * D * 128, &tmW, w_full, kb * 64, 0);
      tc::mbar_wait(w_full, 0);
      const uint32_t idesc = tc::make_idesc_bf16(TILE_M, D, 0, 0);
      const int ksteps = p.PD / 16;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
        const int s = it & 1;
        const uint32_t ph = (uint32_t)((it >> 1) & 1);
        tc::mbar_wait(&acc_empty[s], ph ^ 1);
        tc::mbar_wait(&a_full[s], ph);
        tc::fence_after_sync();
        const uint32_t sa = tc::smem_u32(smem + Smem::OFF_A + s * A_BYTES), sw = tc::smem_u32(smem + Smem::OFF_W);
        for (int k = 0; k < ksteps; ++k) {
          const uint64_t ad = tc::make_sdesc_sw128(sa + (k >> 2) * (TILE_M * 128) + (k & 3) * 32, 16, 1024);
          const uint64_t bd = tc::make_sdesc_sw128(sw + (k >> 2) * (D * 128) + (k & 3) * 32, 16, 1024);
          tc::mma_bf16_ss(tmem_base + s * D, ad, bd, idesc, k > 0 ? 1u : 0u);
        }
        tc::mma_commit(&a_empty[s]);
        tc::mma_commit(&acc_full[s]);
      }
    }
  } else if (warp <= PROD_WARPS) {
    // ===================================== producers: patchify + LayerNorm(patch_dim) =====================================
    const int pw = warp - 1;                              // 0..7
    const ImgT* img = reinterpret_cast<const ImgT*>(p.img);
    // token rows: x0[token row of sample b] = token + pos_embedding[0]
    {
      const int S1 = p.np + 1;
      for (int b = blockIdx.x * PROD_WARPS + pw; b < p.B; b += gridDim.x * PROD_WARPS) {
        float* dst = p.x0 + ((p.mode == 0 ? (size_t)p.B * p.np : 0) + (size_t)b * S1) * D;
        for (int c = lane; c < D / 4; c += 32) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(p.tok) + c), q = __ldg(reinterpret_cast<const float4*>(p.pos) + c);
          reinterpret_cast<float4*>(dst)[c] = make_float4(t.x + q.x, t.y + q.y, t.z + q.z, t.w + q.w);
        }
      }
    }
    // this lane's elements of a patch: k = 2 lane + 64 i + {0, 1}, (p1 p2 c) order (vit.py:110): pixel (c, p1, p2) of the patch
    float g1v[3][2], b1v[3][2];
    int goff[3][2];                                       // offset of the element inside the sample image, relative to the patch origin
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int k = 2 * lane + 64 * i + e;
        const bool ok = k < p.PD;
        g1v[i][e] = ok ? p.g1[k] : 0.f;
        b1v[i][e] = ok ? p.b1[k] : 0.f;
        const int c = k % p.C, p2 = (k / p.C) % p.P, p1 = k / (p.C * p.P);
        goff[i][e] = ok ? (c * p.HW + p1) * p.HW + p2 : -1;
      }
    // The pixels are gathered straight from global memory: a patch is C x P runs of P contiguous pixels, neighbouring patches
    // share their 32-byte sectors (L1 hits), and nothing synchronises the eight producer warps with each other -- the first
    // version staged whole image rows in shared memory behind two CTA-wide barriers per patch row and spent 80 us per tile
    // waiting on them.  Two patches per warp are in flight.
    auto load_patch = [&](int row, float (&v)[3][2]) {
      const int b = row / p.np, j = row - b * p.np, prow = j / p.npr, pcol = j - prow * p.npr;
      const long long src = p.sample_idx ? p.sample_idx[b] : (long long)b;
      const ImgT* base = img + (size_t)src * p.C * p.HW * p.HW + (size_t)(prow * p.P) * p.HW + pcol * p.P;
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int e = 0; e < 2; ++e) v[i][e] = goff[i][e] >= 0 ? pix(base + goff[i][e]) : 0.f;
    };
    // uint8 frames (what the device-resident rollout holds): the C x P image rows of a patch row (2 016 bytes) are staged with
    // cp.async into a ring of four buffers, two patch rows ahead of the one being normalised -- the gather latency is off the
    // critical path and every pixel is read from global memory once, as whole 84-byte rows.
    int soff[3][2];                                       // byte offset of the element inside a staged patch row, relative to the patch
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int k = 2 * lane + 64 * i + e;
        const int c = k % p.C, p2 = (k / p.C) % p.P, p1 = k / (p.C * p.P);
        soff[i][e] = k < p.PD ? (c * p.P + p1) * p.HW + p2 : 0;
      }
    const int ptid = threadIdx.x - 32;                    // 0..255
    const int pieces = (p.C * p.P * p.HW) / 4;            // 4-byte cp.async pieces per patch row (HW % 4 == 0)
    const int ppr = p.HW / 4;                             // pieces per image row
    int it = 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
      const int s = it & 1;
      const uint32_t ph = (uint32_t)((it >> 1) & 1);
      tc::mbar_wait(&a_empty[s], ph ^ 1);
      uint8_t* A = smem + Smem::OFF_A + s * A_BYTES;
      const int row_lo = tile * TILE_M, row_hi = min(p.rows, row_lo + TILE_M);
      auto finish_patch = [&](int row, const float (&v)[3][2]) {
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < 3; ++i) sum += v[i][0] + v[i][1];
        const float mu = warp_sum(sum) / (float)p.PD;
        float qq = 0.f;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int e = 0; e < 2; ++e)
            if (goff[i][e] >= 0) { const float d = v[i][e] - mu; qq += d * d; }
        const float rs = rsqrtf(warp_sum(qq) / (float)p.PD + p.eps1);
        if (lane == 0) { p.pmean[row] = mu; p.prstd[row] = rs; }
        const int tr = row - row_lo;                       // row inside the tile
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          if (goff[i][0] >= 0) {
            const uint32_t pk = pack_bf16x2((v[i][0] - mu) * rs * g1v[i][0] + b1v[i][0], (v[i][1] - mu) * rs * g1v[i][1] + b1v[i][1]);
            // element k of k-block i: byte 4 lane of the 128-byte row = chunk lane / 4, offset (lane & 3) * 4
            *reinterpret_cast<uint32_t*>(A + i * (TILE_M * 128) + sw_off(tr, lane >> 2) + (lane & 3) * 4) = pk;
            *reinterpret_cast<uint32_t*>(p.pln + (size_t)row * p.PD + 2 * lane + 64 * i) =
                p.pln_xhat ? pack_bf16x2((v[i][0] - mu) * rs, (v[i][1] - mu) * rs) : pk;
          }
        }
      };
      if constexpr (sizeof(ImgT) == 1) {
        // patch rows that intersect the tile: (first patch row index r, sample, patch row, first patch column, count)
        auto prow_of = [&](int r, int& b, int& prow, int& pc0, int& n_here) {
          b = r / p.np;
          const int j = r - b * p.np;
          prow = j / p.npr;
          pc0 = j - prow * p.npr;
          n_here = min(p.npr - pc0, row_hi - r);
        };
        auto prefetch = [&](int r, int slot) {              // all 256 producer threads
          if (r < row_hi) {
            int b, prow, pc0, n_here;
            prow_of(r, b, prow, pc0, n_here);
            const long long src = p.sample_idx ? p.sample_idx[b] : (long long)b;
            const uint8_t* base = reinterpret_cast<const uint8_t*>(p.img) + (size_t)src * p.C * p.HW * p.HW + (size_t)prow * p.P * p.HW;
            const uint32_t dst = tc::smem_u32(smem + Smem::OFF_STAGE + slot * STAGE_BYTES);
            for (int i = ptid; i < pieces; i += PROD_WARPS * 32) {
              const int cr = i / ppr, xw = i - cr * ppr;     // cr = c * P + p1
              const int c = cr / p.P, p1 = cr - c * p.P;
              const uint8_t* g = base + ((size_t)c * p.HW + p1) * p.HW + 4 * xw;
              asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + 4 * i), "l"(g) : "memory");
            }
          }
          asm volatile("cp.async.commit_group;" ::: "memory");   // always one group per call: the wait counts groups
        };
        auto advance = [&](int r) {                         // first patch row index of the NEXT patch row
          int b, prow, pc0, n_here;
          prow_of(r, b, prow, pc0, n_here);
          return r + n_here;
        };
        int r0 = row_lo, r1 = row_lo < row_hi ? advance(row_lo) : row_hi, r2 = r1 < row_hi ? advance(r1) : row_hi;
        int slot = 0;
        prefetch(r0, 0);
        prefetch(r1, 1);
        while (r0 < row_hi) {
          prefetch(r2, (slot + 2) % NSTAGE);                // two patch rows ahead; its buffer was consumed two barriers ago
          asm volatile("cp.async.wait_group 2;" ::: "memory");
          prod_bar();                                       // every producer thread's pieces of patch row r0 have landed
          int b, prow, pc0, n_here;
          prow_of(r0, b, prow, pc0, n_here);
          const uint8_t* st8 = smem + Smem::OFF_STAGE + slot * STAGE_BYTES;
          for (int q = pw; q < n_here; q += PROD_WARPS) {
            const uint8_t* tp = st8 + (pc0 + q) * p.P;
            float v[3][2];
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
              for (int e = 0; e < 2; ++e) v[i][e] = goff[i][e] >= 0 ? s_lut[tp[soff[i][e]]] : 0.f;
            finish_patch(r0 + q, v);
          }
          r0 = r1; r1 = r2; r2 = r2 < row_hi ? advance(r2) : row_hi;
          slot = (slot + 1) % NSTAGE;
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      } else {
        for (int row = row_lo + pw; row < row_hi; row += 4 * PROD_WARPS) {     // four patches per warp in flight
          float va[3][2], vb[3][2], vc[3][2], vd[3][2];
          const bool hb = row + PROD_WARPS < row_hi, hc = row + 2 * PROD_WARPS < row_hi, hd = row + 3 * PROD_WARPS < row_hi;
          load_patch(row, va);
          if (hb) load_patch(row + PROD_WARPS, vb);
          if (hc) load_patch(row + 2 * PROD_WARPS, vc);
          if (hd) load_patch(row + 3 * PROD_WARPS, vd);
          finish_patch(row, va);
          if (hb) finish_patch(row + PROD_WARPS, vb);
          if (hc) finish_patch(row + 2 * PROD_WARPS, vc);
          if (hd) finish_patch(row + 3 * PROD_WARPS, vd);
        }
      }
      // rows past the end of the last tile keep whatever the buffer held: finite, never stored
      tc::fence_proxy_async();                             // generic-proxy writes of A -> visible to the tensor core
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&a_full[s]);
    }
  } else {
    // ===================================== epilogue: + bias, LayerNorm(dim), token / position assembly =====================================
    const int q = warp & 3;                                // TMEM lane quadrant this warp may read
    float* tile_s = reinterpret_cast<float*>(smem + Smem::OFF_EPI) + (warp - 1 - PROD_WARPS) * 1024;
    const int cchunk = lane & 7, rsub = lane >> 3;
    const int S1 = p.np + 1;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
      const int s = it & 1;
      const uint32_t ph = (uint32_t)((it >> 1) & 1);
      tc::mbar_wait(&acc_full[s], ph);
      tc::fence_after_sync();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * D);
      const int row0 = tile * TILE_M + q * 32;             // first patch row of this warp
      const int my_row = row0 + lane;
      // rows this lane stores in the transposed layout: rl = 4 i8 + rsub -> destination rows and positional row, once per tile
      int dstB[8], posr[8];
#pragma unroll
      for (int i8 = 0; i8 < 8; ++i8) {
        const int grow = row0 + i8 * 4 + rsub;
        const int b = grow / p.np, j = grow - b * p.np;
        posr[i8] = grow < p.rows ? 1 + j : -1;
        dstB[i8] = (p.mode == 0 ? p.B * p.np : 0) + b * S1 + 1 + j;
      }
      // ---- pass 1: row statistics of v = acc + bias (this thread owns the whole row: no cross-thread reduction)
      float sum = 0.f, sq = 0.f;
#pragma unroll 1
      for (int c = 0; c < D / 32; ++c) {
        uint32_t rr[32];
        tc::tmem_ld_32x32(taddr + c * 32, rr);
        tc::tmem_ld_wait();
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 bb = *reinterpret_cast<const float4*>(s_bias + c * 32 + j4 * 4);       // broadcast read
          const float v0 = __uint_as_float(rr[4 * j4]) + bb.x, v1 = __uint_as_float(rr[4 * j4 + 1]) + bb.y;
          const float v2 = __uint_as_float(rr[4 * j4 + 2]) + bb.z, v3 = __uint_as_float(rr[4 * j4 + 3]) + bb.w;
          sum += (v0 + v1) + (v2 + v3);
          sq += fmaf(v0, v0, v1 * v1) + fmaf(v2, v2, v3 * v3);
        }
      }
      const float mean = sum * (1.0f / D);
      const float rstd = rsqrtf(fmaxf(sq * (1.0f / D) - mean * mean, 0.f) + p.eps3);
      if (my_row < p.rows) { p.m3[my_row] = mean; p.r3[my_row] = rstd; }
      // statistics of the rows this lane handles after the transpose (row rl = 4 i8 + rsub lives in lane rl of this warp)
      float ra[8], rb[8];
#pragma unroll
      for (int i8 = 0; i8 < 8; ++i8) {
        const float m_ = __shfl_sync(0xffffffffu, mean, i8 * 4 + rsub), r_ = __shfl_sync(0xffffffffu, rstd, i8 * 4 + rsub);
        ra[i8] = r_;
        rb[i8] = -m_ * r_;
      }
      // ---- pass 2: the raw accumulator goes through the transpose tile once; in the row-contiguous layout a lane adds the bias,
      // stores e0 (the LayerNorm input, kept for the backward), normalises, and stores the token into the explorative row as is
      // and + pos_embedding[1 + j] into the exploitative / CLS row
#pragma unroll 1
      for (int c = 0; c < D / 32; ++c) {
        const int col = c * 32 + cchunk * 4;
        const float4 bb = *reinterpret_cast<const float4*>(s_bias + col);
        const float4 gg = *reinterpret_cast<const float4*>(s_g3 + col);
        const float4 be = *reinterpret_cast<const float4*>(s_b3 + col);
        float4 pe[8];                                       // positional rows of this chunk: in flight under the TMEM load
#pragma unroll
        for (int i8 = 0; i8 < 8; ++i8)
          if (posr[i8] >= 0) pe[i8] = __ldg(reinterpret_cast<const float4*>(p.pos + (size_t)posr[i8] * D + col));
        uint32_t rr[32];
        tc::tmem_ld_32x32(taddr + c * 32, rr);
        tc::tmem_ld_wait();
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4)
          *reinterpret_cast<uint4*>(tile_s + lane * 32 + ((j4 ^ (lane & 7)) << 2)) = make_uint4(rr[4 * j4], rr[4 * j4 + 1], rr[4 * j4 + 2], rr[4 * j4 + 3]);
        __syncwarp();
#pragma unroll
        for (int i8 = 0; i8 < 8; ++i8) {
          const int rl = i8 * 4 + rsub;
          if (posr[i8] >= 0) {
            float4 v = *reinterpret_cast<const float4*>(tile_s + rl * 32 + ((cchunk ^ (rl & 7)) << 2));
            v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
            *reinterpret_cast<float4*>(p.e0 + (size_t)(row0 + rl) * D + col) = v;
            const float4 t = make_float4(fmaf(fmaf(v.x, ra[i8], rb[i8]), gg.x, be.x), fmaf(fmaf(v.y, ra[i8], rb[i8]), gg.y, be.y),
                                         fmaf(fmaf(v.z, ra[i8], rb[i8]), gg.z, be.z), fmaf(fmaf(v.w, ra[i8], rb[i8]), gg.w, be.w));
            if (p.mode == 0) *reinterpret_cast<float4*>(p.x0 + (size_t)(row0 + rl) * D + col) = t;          // no pos-emb (bug kept)
            *reinterpret_cast<float4*>(p.x0 + (size_t)dstB[i8] * D + col) =
                make_float4(t.x + pe[i8].x, t.y + pe[i8].y, t.z + pe[i8].z, t.w + pe[i8].w);
          }
        }
        __syncwarp();
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&acc_empty[s]);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace ef
}  // namespace eavit

using namespace eavit;

extern "C" int eavit_embed_fused_fwd(const void* img, int img_dtype, const long long* sample_idx, int B, int C, int HW, int P, int mode,
                                     const float* g1, const float* b1, float eps1, const void* w_bf16, const float* bias,
                                     const float* g3, const float* b3, float eps3, const float* pos, const float* tok,
                                     void* pln_bf16, float* pmean, float* prstd, float* e0, float* m3, float* r3, float* x0,
                                     int pln_xhat, void* stream) {
  EAVIT_CHECK_ARG(img && g1 && b1 && w_bf16 && bias && g3 && b3 && pos && tok && pln_bf16 && pmean && prstd && e0 && m3 && r3 && x0);
  EAVIT_CHECK_ARG(B > 0 && C > 0 && P > 0 && HW % P == 0 && (mode == 0 || mode == 1));
  EAVIT_CHECK_ARG(img_dtype == EAVIT_U8 || img_dtype == EAVIT_F32);
  const int PD = C * P * P, npr = HW / P, np = npr * npr;
  EAVIT_CHECK_ARG(PD % 16 == 0 && PD <= ef::KPAD && HW % 4 == 0 && C * P * HW <= ef::STAGE_BYTES);
  EAVIT_CHECK_ARG((long long)B * (np + 1) * 2 < (1LL << 31));
  EAVIT_CHECK_ARG((reinterpret_cast<uintptr_t>(w_bf16) & 15) == 0 && (PD * 2) % 16 == 0);
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap tmW;
  int rc = make_tmap_bf16_2d(&tmW, w_bf16, (uint64_t)PD, (uint64_t)ef::D, (uint64_t)PD * 2, ef::D);
  if (rc) return rc;
  ef::Params p;
  p.img = img; p.img_u8 = img_dtype == EAVIT_U8; p.sample_idx = sample_idx;
  p.B = B; p.C = C; p.HW = HW; p.P = P; p.PD = PD; p.np = np; p.npr = npr; p.mode = mode;
  p.g1 = g1; p.b1 = b1; p.eps1 = eps1; p.bias = bias; p.g3 = g3; p.b3 = b3; p.eps3 = eps3; p.pos = pos; p.tok = tok;
  p.pln = reinterpret_cast<__nv_bfloat16*>(pln_bf16); p.pmean = pmean; p.prstd = prstd; p.e0 = e0; p.m3 = m3; p.r3 = r3; p.x0 = x0;
  p.rows = B * np;
  p.pln_xhat = pln_xhat;
  p.tiles = cdiv(p.rows, ef::TILE_M);
  const int grid = p.tiles < kNumSMs ? p.tiles : kNumSMs;
  static bool attr_done = false;
  if (!attr_done) {
    EAVIT_CUDA(cudaFuncSetAttribute(ef::embed_fused_kernel<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, ef::Smem::TOTAL));
    EAVIT_CUDA(cudaFuncSetAttribute(ef::embed_fused_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, ef::Smem::TOTAL));
    attr_done = true;
  }
  if (p.img_u8) ef::embed_fused_kernel<uint8_t><<<grid, ef::THREADS, ef::Smem::TOTAL, st>>>(tmW, p);
  else ef::embed_fused_kernel<float><<<grid, ef::THREADS, ef::Smem::TOTAL, st>>>(tmW, p);
  EAVIT_LAUNCH_OK();
  return EAVIT_OK;
}
