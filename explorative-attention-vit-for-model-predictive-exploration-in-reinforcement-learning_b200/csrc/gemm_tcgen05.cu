// Persistent, warp-specialised bf16 GEMM on tcgen05 tensor cores with TMEM accumulators and a fused epilogue.
//
//   C[M,N] = epilogue( A . B^T )      A: [M,K] (K-major) or [K,M] (MN-major),  B: [N,K] or [K,N]
//
// One CTA per SM, 576 threads:
//   warp 0      TMA producer   (cp.async.bulk.tensor, 128-byte swizzle, STAGES-deep mbarrier ring)
//   warp 1      TMEM allocator (2 accumulator stages x BN columns) + MMA issuer (one lane issues tcgen05.mma
//               128 x BN x 16 and commits to mbarriers)
//   warps 2-17  epilogue       (tcgen05.ld -> bias / activation / residual -> 16-byte global stores); four warps per
//               TMEM lane quadrant: at K = 256 the fused epilogues (GELU, GELU', residual) are 2-3x longer than the
//               mainloop and latency-bound, so they need >= 4 warps per scheduler to keep the issue slots busy
// The accumulator is double-buffered in TMEM so the epilogue of tile i overlaps the mainloop of tile
// i+1 -- at K = 256 (the ViT width) the epilogue is as long as the mainloop, so this overlap is what
// keeps the tensor pipe busy.  Tiles are scheduled n-fastest so CTAs that share an A row-block run
// together and re-read it from L2, not HBM.
//
// Replaces: every nn.Linear on the ViT path (reference vit.py:29-32, :52-57, :112), the HF q/k/v/dense
// layers (vit_hg.py via transformers), the RND fully-connected layers and im2col'ed convolutions
// (model.py:368-416), and all of their backward GEMMs (dX = dY W, dW = dY^T X via MN-major operands).
#include "common.cuh"
#include "tc05.cuh"

namespace eavit {

constexpr int BM = 128;
constexpr int BK = 64;                 // 64 bf16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
constexpr int CTRL_WARPS = 2;
constexpr int MAX_EPI_WARPS = 16;
constexpr int EPI_TILE_FLOATS = 32 * 32;   // 32x32 transpose tile per epilogue warp (16-byte chunks XOR-swizzled)
constexpr int A_STAGE_BYTES = BM * BK * 2;

struct GemmKernelParams {
  int M, N, K;
  int a_mn, b_mn;
  int m_tiles, n_tiles, splits, kb_per_split, kb_total;
  const float* bias;
  const __nv_bfloat16* aux;
  const float* residual;
  float* out_f32;
  __nv_bfloat16* out_bf16;
  __nv_bfloat16* out_pre;
  float* colsum;
  long long ldc;
  DropCfg drop;
  int act;
  int atomic_f32;
};

template <int BN>
struct GemmCfg {
  static constexpr int B_STAGE_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int STAGES = (BN == 256) ? 3 : (BN == 128 ? 5 : 6);
  static constexpr int TMEM_COLS = 2 * BN;    // power of two >= 32 for BN in {64,128,256}
  static constexpr int EPI_BYTES = MAX_EPI_WARPS * EPI_TILE_FLOATS * 4;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

// ---------------------------------------------------------------------------------------------------- GELU by table
// The GELU epilogues were instruction-issue bound (ncu r1: 53-66 % issue-active, ~30 instructions and 2 MUFU per element,
// 180 us of pure issue time for the [201216, 1024] MLP activations against 143 us of HBM time).  Both take their argument
// as a BFLOAT16 value -- the stored pre-activation `hpre` -- and a bf16 has only 2^16 bit patterns: with |x| clamped to
// [2^-14, 8) (below: Phi = 0.5 +- 1e-5, gelu' = 0.5 +- 4e-5; above: 0 / 1 to fp32 accuracy) 2 x 2176 entries cover every
// input.  Each CTA builds the 17 KB fp32 table in its prologue (normcdff / expf, under the previous kernel's tail) and
// the epilogue does one shared-memory lookup per element:
//   forward   h = v * Phi(bf16(v))            v = fp32 accumulator + bias; bf16(v) is the pre-activation that is stored anyway
//   backward  dh = dy * gelu'(hpre)           gelu'(x) = Phi(x) + x phi(x)
// Evaluating Phi at the rounded argument costs |v| phi(v) |v| 2^-9 <= 6e-4 absolute on h, a quarter of the bf16 rounding of
// h itself (vit.py:30 nn.GELU(): exact-erf GELU).
constexpr int GELU_LUT_LO = 0x3880;                 // bf16 bits of 2^-14
constexpr int GELU_LUT_HI = 0x4100;                 // bf16 bits of 8.0
constexpr int GELU_LUT_N = GELU_LUT_HI - GELU_LUT_LO;       // 2176 magnitudes per sign
constexpr int GELU_LUT_BYTES = 2 * GELU_LUT_N * 4;

template <bool GRAD>
__device__ __forceinline__ void gelu_lut_build(float* lut, int tid, int nthreads) {
  for (int i = tid; i < 2 * GELU_LUT_N; i += nthreads) {
    const int sgn = i >= GELU_LUT_N, a = (sgn ? i - GELU_LUT_N : i) + GELU_LUT_LO;
    const float x = __uint_as_float(((uint32_t)(sgn << 15 | a)) << 16);
    const float cdf = normcdff(x);
    lut[i] = GRAD ? fmaf(x * 0.39894228040143268f, __expf(-0.5f * x * x), cdf) : cdf;
  }
}
// table entry for the bf16 value whose bits are the low 16 bits of `b`
__device__ __forceinline__ float gelu_lut(const float* lut, uint32_t b) {
  const int a = min(max((int)(b & 0x7fffu), GELU_LUT_LO), GELU_LUT_HI - 1) - GELU_LUT_LO;
  return lut[a + (int)((b >> 15) & 1u) * GELU_LUT_N];
}

__device__ __forceinline__ float apply_act(float v, int act, float a) {
  switch (act) {
    case EAVIT_ACT_GELU: return gelu_erf(v);
    case EAVIT_ACT_GELU_BWD: return v * gelu_erf_grad(a);
    case EAVIT_ACT_LRELU: return v > 0.f ? v : 0.01f * v;
    case EAVIT_ACT_LRELU_BWD: return a > 0.f ? v : 0.01f * v;
    case EAVIT_ACT_RELU: return fmaxf(v, 0.f);
    case EAVIT_ACT_RELU_BWD: return a > 0.f ? v : 0.f;
    default: return v;
  }
}

// Epilogue specialisations (compile-time): the epilogue is instruction-issue bound (ncu: 45 % issue-active from only 8
// warps), so the common fused forms drop every per-element runtime branch of the generic path.
enum { E_GENERIC = 0, E_STORE = 1, E_GELU_FWD = 2, E_GELU_BWD = 3, E_RESID = 4, E_ATOMIC = 5 };
// The GELU epilogues are issue/latency-bound (16 instructions + 2 MUFU per element): 16 warps.  The store / residual /
// atomic epilogues are HBM-bound and want registers for loads in flight instead: 8 warps.
template <int EPI> struct EpiWarps { static constexpr int N = (EPI == E_GELU_FWD || EPI == E_STORE) ? 16 : (EPI == E_GELU_BWD ? 12 : 8); };

template <int EPI> struct EpiLut { static constexpr int BYTES = (EPI == E_GELU_FWD || EPI == E_GELU_BWD) ? GELU_LUT_BYTES : 0; };

template <int BN, int EPI, bool DROP>      // DROP: dropout mask in the epilogue (compile-time: the branch costs the fused epilogues 6-17 %)
__global__ void __launch_bounds__((CTRL_WARPS + EpiWarps<EPI>::N) * 32, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const GemmKernelParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int EPI_WARPS = EpiWarps<EPI>::N;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* epi_smem = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES + Cfg::EPI_BYTES);
  uint64_t* full_bar = bars;                      // [STAGES]   TMA -> MMA
  uint64_t* empty_bar = bars + STAGES;            // [STAGES]   MMA -> TMA
  uint64_t* acc_full = bars + 2 * STAGES;         // [2]        MMA -> epilogue
  uint64_t* acc_empty = bars + 2 * STAGES + 2;    // [2]        epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  float* gelu_tab = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES + Cfg::EPI_BYTES + 256);   // behind the barrier block
  if constexpr (EPI == E_GELU_FWD) gelu_lut_build<false>(gelu_tab, threadIdx.x, blockDim.x);
  if constexpr (EPI == E_GELU_BWD) gelu_lut_build<true>(gelu_tab, threadIdx.x, blockDim.x);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = p.m_tiles * p.n_tiles * p.splits;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmA);
    tc::prefetch_tmap(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) { tc::mbar_init(&full_bar[i], 1); tc::mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&acc_full[i], 1); tc::mbar_init(&acc_empty[i], EPI_WARPS); }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: everything above (barriers, TMEM allocation, descriptor prefetch) may run while the
  // previous kernel of the stream drains; its results are only touched below this wait.  The next kernel is released
  // for its own prologue right away -- it cannot become resident on an SM before this CTA's shared memory is gone.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n_blk = tile % p.n_tiles, m_blk = (tile / p.n_tiles) % p.m_tiles, split = tile / (p.n_tiles * p.m_tiles);
        const int kb0 = split * p.kb_per_split, kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          tc::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + A_STAGE_BYTES;
          tc::mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          if (!p.a_mn) {
            tc::tma_load_2d(sa, &tmA, &full_bar[stage], kb * BK, m_blk * BM);
          } else {
#pragma unroll
            for (int i = 0; i < BM / 64; ++i)
              tc::tma_load_2d(sa + i * 8192, &tmA, &full_bar[stage], m_blk * BM + i * 64, kb * BK);
          }
          if (!p.b_mn) {
            tc::tma_load_2d(sb, &tmB, &full_bar[stage], kb * BK, n_blk * BN);
          } else {
#pragma unroll
            for (int i = 0; i < BN / 64; ++i)
              tc::tma_load_2d(sb + i * 8192, &tmB, &full_bar[stage], n_blk * BN + i * 64, kb * BK);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =====================================
    if (lane == 0) {
      const uint32_t idesc = tc::make_idesc_bf16(BM, BN, p.a_mn, p.b_mn);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int split = tile / (p.n_tiles * p.m_tiles);
        const int kb0 = split * p.kb_per_split, kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        tc::mbar_wait(&acc_empty[acc], acc_phase ^ 1);
        tc::fence_after_sync();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          tc::mbar_wait(&full_bar[stage], phase);
          tc::fence_after_sync();
          const uint32_t sa = tc::smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t sb = sa + A_STAGE_BYTES;
          // K-major:  rows of 128 B, 8-row groups 1024 B apart (SBO); K advance = 32 B inside the swizzle atom.
          // MN-major: 64-element (128 B) MN chunks, 8 k-rows per 1024 B (SBO), MN blocks 8192 B apart (LBO);
          //           K advance = 16 k-rows = 2048 B.
          const uint64_t a_desc = p.a_mn ? tc::make_sdesc_sw128(sa, 8192, 1024) : tc::make_sdesc_sw128(sa, 16, 1024);
          const uint64_t b_desc = p.b_mn ? tc::make_sdesc_sw128(sb, 8192, 1024) : tc::make_sdesc_sw128(sb, 16, 1024);
          const uint32_t a_step = p.a_mn ? (2048 >> 4) : (32 >> 4);
          const uint32_t b_step = p.b_mn ? (2048 >> 4) : (32 >> 4);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            tc::mma_bf16_ss(d_tmem, a_desc + (uint64_t)(k * a_step), b_desc + (uint64_t)(k * b_step), idesc,
                            (kb > kb0 || k > 0) ? 1u : 0u);
          }
          tc::mma_commit(&empty_bar[stage]);          // frees the smem slot when these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tc::mma_commit(&acc_full[acc]);               // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================================== epilogue =====================================
    // Four warps per TMEM lane quadrant (they interleave 32-column chunks).  Each chunk is transposed through a padded
    // shared-memory tile so that global loads / stores are row-contiguous (a warp touches 128 B of ONE row per
    // instruction instead of 32 rows x 16 B).
    const int e = warp - CTRL_WARPS;
    const int q = warp & 3;                           // TMEM lane quadrant this warp may read (hardware: warp id % 4)
    const int par = e >> 2;                           // first 32-column chunk handled by this warp
    float* tile = epi_smem + e * EPI_TILE_FLOATS;
    constexpr bool GEN = EPI == E_GENERIC;
    const bool need_aux = GEN ? (p.act == EAVIT_ACT_GELU_BWD || p.act == EAVIT_ACT_LRELU_BWD || p.act == EAVIT_ACT_RELU_BWD) : (EPI == E_GELU_BWD);
    const bool has_bias = GEN || EPI == E_STORE ? (p.bias != nullptr) : (EPI == E_GELU_FWD || EPI == E_RESID);
    const bool has_pre = GEN ? (p.out_pre != nullptr) : (EPI == E_GELU_FWD);
    const bool has_res = GEN ? (p.residual != nullptr) : (EPI == E_RESID);
    const bool out_f32 = GEN || EPI == E_STORE ? (p.out_f32 != nullptr) : (EPI == E_RESID || EPI == E_ATOMIC);
    const bool out_b16 = GEN || EPI == E_STORE ? (p.out_bf16 != nullptr) : (EPI == E_GELU_FWD || EPI == E_GELU_BWD);
    const bool atomic = GEN ? (p.atomic_f32 != 0) : (EPI == E_ATOMIC);
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile_i = blockIdx.x; tile_i < total_tiles; tile_i += gridDim.x) {
      const int n_blk = tile_i % p.n_tiles, m_blk = (tile_i / p.n_tiles) % p.m_tiles;
      tc::mbar_wait(&acc_full[acc], acc_phase);
      tc::fence_after_sync();
      const int row0 = m_blk * BM + q * 32;
      const int nrows = min(32, p.M - row0);          // may be <= 0 for the M tail
#pragma unroll 1
      for (int c = par; c < BN / 32; c += EPI_WARPS / 4) {
        const int col0 = n_blk * BN + c * 32;
        if (col0 >= p.N) break;                        // warp-uniform
        const int cchunk = lane & 7, rsub = lane >> 3;
        const int col = col0 + cchunk * 4;
        constexpr bool CS = GEN || EPI == E_GELU_BWD;    // epilogues that can carry a fused column sum
        const size_t off0 = (size_t)(row0 + rsub) * (size_t)p.ldc + col;      // row rr = rsub + 4 * it
        const size_t ostep = 4 * (size_t)p.ldc;
        // Issue every global read of this chunk FIRST (8 independent 16-byte loads per lane: the epilogue is latency-
        // bound on HBM unless ~40 KB per SM are in flight) -- before the TMEM load and the __syncwarp of the transpose,
        // which also stops the scheduler from sinking them between the stores further down.
        float4 res[8];
        uint2 ax[8];
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        auto load = [&](int it) {
          if (has_res) res[it] = __ldg(reinterpret_cast<const float4*>(p.residual + off0 + it * ostep));
          if (need_aux) ax[it] = __ldg(reinterpret_cast<const uint2*>(p.aux + off0 + it * ostep));
        };
        if (col < p.N) {
          if (has_bias) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
          if (has_res || need_aux) {
            if (nrows >= 32) {
#pragma unroll
              for (int it = 0; it < 8; ++it) load(it);
            } else {
#pragma unroll
              for (int it = 0; it < 8; ++it)
                if (it * 4 + rsub < nrows) load(it);
            }
          }
        }
        // registers (one row per lane) -> swizzled smem tile -> (4 rows x 8 lanes x 16 B) per instruction; two 16-column
        // halves so that only 16 accumulator registers are live next to the prefetched residual / aux values
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t r[16];
          tc::tmem_ld_32x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c * 32 + hh * 16), r);
          tc::tmem_ld_wait();
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4)
            *reinterpret_cast<uint4*>(tile + lane * 32 + (((hh * 4 + c4) ^ (lane & 7)) << 2)) = make_uint4(r[4 * c4], r[4 * c4 + 1], r[4 * c4 + 2], r[4 * c4 + 3]);
        }
        __syncwarp();
        float cs[4] = {0.f, 0.f, 0.f, 0.f};              // column sums of the stored values (bias gradient)
        if (col < p.N) {
          auto body = [&](int it) {
            const int rr = it * 4 + rsub;
            const float4 t = *reinterpret_cast<const float4*>(tile + rr * 32 + ((cchunk ^ (rr & 7)) << 2));
            float v[4] = {t.x + b4.x, t.y + b4.y, t.z + b4.z, t.w + b4.w};
            const size_t off = off0 + it * ostep;
            if constexpr (EPI == E_GELU_FWD) {
              const uint32_t p01 = pack_bf16x2(v[0], v[1]), p23 = pack_bf16x2(v[2], v[3]);      // the stored pre-activation
              *reinterpret_cast<uint2*>(p.out_pre + off) = make_uint2(p01, p23);
              v[0] *= gelu_lut(gelu_tab, p01); v[1] *= gelu_lut(gelu_tab, p01 >> 16);
              v[2] *= gelu_lut(gelu_tab, p23); v[3] *= gelu_lut(gelu_tab, p23 >> 16);
            } else if (has_pre) {
              *reinterpret_cast<uint2*>(p.out_pre + off) = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
            }
            if constexpr (EPI == E_GELU_BWD) {
              v[0] *= gelu_lut(gelu_tab, ax[it].x); v[1] *= gelu_lut(gelu_tab, ax[it].x >> 16);
              v[2] *= gelu_lut(gelu_tab, ax[it].y); v[3] *= gelu_lut(gelu_tab, ax[it].y >> 16);
            } else if constexpr (GEN) {
              if (p.act != EAVIT_ACT_NONE) {
                float a[4] = {0.f, 0.f, 0.f, 0.f};
                if (need_aux) {
                  const float2 lo = unpack_bf16x2(ax[it].x), hi = unpack_bf16x2(ax[it].y);
                  a[0] = lo.x; a[1] = lo.y; a[2] = hi.x; a[3] = hi.y;
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) v[i] = apply_act(v[i], p.act, a[i]);
              }
            }
            if constexpr (DROP) {                     // nn.Dropout after the Linear / activation
              float dm[4];
              drop4(p.drop, drop_row_key(p.drop, (uint32_t)(row0 + rr)), (uint32_t)col, dm);
              v[0] *= dm[0]; v[1] *= dm[1]; v[2] *= dm[2]; v[3] *= dm[3];
            }
            if (has_res) { v[0] += res[it].x; v[1] += res[it].y; v[2] += res[it].z; v[3] += res[it].w; }
            if constexpr (CS) { cs[0] += v[0]; cs[1] += v[1]; cs[2] += v[2]; cs[3] += v[3]; }
            if (out_f32) {
              if (atomic) {
#pragma unroll
                for (int i = 0; i < 4; ++i) atomicAdd(p.out_f32 + off + i, v[i]);
              } else {
                *reinterpret_cast<float4*>(p.out_f32 + off) = make_float4(v[0], v[1], v[2], v[3]);
              }
            }
            if (out_b16)
              *reinterpret_cast<uint2*>(p.out_bf16 + off) = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
          };
          if (nrows >= 32) {                               // full tile rows (all but the last M tile): no per-row predicates
#pragma unroll
            for (int it = 0; it < 8; ++it) body(it);
          } else {
#pragma unroll
            for (int it = 0; it < 8; ++it)
              if (it * 4 + rsub < nrows) body(it);
          }
        }
        if (p.colsum != nullptr) {                         // warp-uniform
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            cs[i] += __shfl_xor_sync(0xffffffffu, cs[i], 8);
            cs[i] += __shfl_xor_sync(0xffffffffu, cs[i], 16);
          }
          if (rsub == 0 && col < p.N) {
#pragma unroll
            for (int i = 0; i < 4; ++i) atomicAdd(p.colsum + col + i, cs[i]);
          }
        }
        __syncwarp();
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&acc_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess || p == nullptr) {
      set_error("cudaGetDriverEntryPoint(cuTensorMapEncodeTiled) failed");
      return nullptr;
    }
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D bf16 tensor map, 128-byte swizzle: dims {inner, outer}, box {64, box_outer}
int make_tmap_bf16_2d(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_bytes,
                      uint32_t box_outer) {
  EncodeTiledFn enc = get_encode_fn();
  if (enc == nullptr) return EAVIT_ECUDA;
  // cuTensorMapEncodeTiled is a DRIVER call: it needs the primary context current on the calling thread.  Threads that
  // have only been handed work (torch's autograd worker threads) may not have touched the runtime yet.
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) { cudaFree(nullptr); ctx_bound = true; }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_bytes};
  cuuint32_t box[2] = {64, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: %d (base %p inner %llu outer %llu pitch %llu box %u)", (int)r, base,
              (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)pitch_bytes, box_outer);
    return EAVIT_ECUDA;
  }
  return EAVIT_OK;
}

template <int BN, int EPI, bool DROP = false>
static int launch_gemm(const eavit_gemm_args* a, cudaStream_t st) {
  using Cfg = GemmCfg<BN>;
  static bool attr_done = false;
  if (!attr_done) {
    EAVIT_CUDA(cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<BN, EPI, DROP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    Cfg::SMEM_BYTES + EpiLut<EPI>::BYTES));
    attr_done = true;
  }
  CUtensorMap tmA, tmB;
  int rc;
  if (!a->a_mn) rc = make_tmap_bf16_2d(&tmA, a->A, (uint64_t)a->K, (uint64_t)a->M, (uint64_t)a->lda * 2, BM);
  else          rc = make_tmap_bf16_2d(&tmA, a->A, (uint64_t)a->M, (uint64_t)a->K, (uint64_t)a->lda * 2, 64);
  if (rc) return rc;
  if (!a->b_mn) rc = make_tmap_bf16_2d(&tmB, a->B, (uint64_t)a->K, (uint64_t)a->N, (uint64_t)a->ldb * 2, BN);
  else          rc = make_tmap_bf16_2d(&tmB, a->B, (uint64_t)a->N, (uint64_t)a->K, (uint64_t)a->ldb * 2, 64);
  if (rc) return rc;

  GemmKernelParams p;
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.a_mn = a->a_mn; p.b_mn = a->b_mn;
  p.m_tiles = cdiv(a->M, BM);
  p.n_tiles = cdiv(a->N, BN);
  p.kb_total = cdiv(a->K, BK);
  int splits = a->split_k > 0 ? a->split_k : 1;
  if (splits > p.kb_total) splits = p.kb_total;
  p.kb_per_split = cdiv(p.kb_total, splits);
  p.splits = cdiv(p.kb_total, p.kb_per_split);
  p.bias = a->bias;
  p.aux = reinterpret_cast<const __nv_bfloat16*>(a->aux_bf16);
  p.residual = a->residual;
  p.out_f32 = a->out_f32;
  p.out_bf16 = reinterpret_cast<__nv_bfloat16*>(a->out_bf16);
  p.out_pre = reinterpret_cast<__nv_bfloat16*>(a->out_pre_bf16);
  p.colsum = a->colsum;
  p.drop = make_drop(a->drop_p, a->drop_seed);
  p.ldc = a->ldc;
  p.act = a->act;
  p.atomic_f32 = a->atomic_f32;
  const int total = p.m_tiles * p.n_tiles * p.splits;
  const int grid = total < kNumSMs ? total : kNumSMs;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3((CTRL_WARPS + EpiWarps<EPI>::N) * 32);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES + EpiLut<EPI>::BYTES;
  static_assert(Cfg::SMEM_BYTES + EpiLut<EPI>::BYTES <= 232448, "over the 227 KB shared-memory limit of a CTA");
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  EAVIT_CUDA(cudaLaunchKernelEx(&cfg, gemm_bf16_tcgen05_kernel<BN, EPI, DROP>, tmA, tmB, p));
  EAVIT_LAUNCH_OK();
  return EAVIT_OK;
}

}  // namespace eavit

using namespace eavit;

extern "C" int eavit_gemm_bf16(const eavit_gemm_args* a, void* stream) {
  EAVIT_CHECK_ARG(a != nullptr);
  EAVIT_CHECK_ARG(a->M > 0 && a->N > 0 && a->K > 0);
  EAVIT_CHECK_ARG(a->A != nullptr && a->B != nullptr);
  EAVIT_CHECK_ARG(a->N % 8 == 0);
  EAVIT_CHECK_ARG((a->lda * 2) % 16 == 0 && (a->ldb * 2) % 16 == 0 && a->ldc % 8 == 0);
  EAVIT_CHECK_ARG((reinterpret_cast<uintptr_t>(a->A) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->B) & 15) == 0);
  EAVIT_CHECK_ARG(a->out_f32 != nullptr || a->out_bf16 != nullptr || a->out_pre_bf16 != nullptr);
  EAVIT_CHECK_ARG(a->split_k <= 1 || (a->atomic_f32 && a->out_f32 != nullptr && a->out_bf16 == nullptr &&
                                      a->out_pre_bf16 == nullptr && a->act == EAVIT_ACT_NONE && a->residual == nullptr &&
                                      a->colsum == nullptr && a->drop_p == 0.f));
  EAVIT_CHECK_ARG(a->drop_p >= 0.f && a->drop_p < 1.f);
  const bool need_aux = (a->act == EAVIT_ACT_GELU_BWD || a->act == EAVIT_ACT_LRELU_BWD || a->act == EAVIT_ACT_RELU_BWD);
  EAVIT_CHECK_ARG(!need_aux || a->aux_bf16 != nullptr);
  EAVIT_CHECK_ARG(a->bias == nullptr || (reinterpret_cast<uintptr_t>(a->bias) & 15) == 0);
  cudaStream_t st = (cudaStream_t)stream;
  const bool drop = make_drop(a->drop_p, a->drop_seed).thresh != 0;
  if (a->N > 128) {
    const bool none = a->act == EAVIT_ACT_NONE;
    const bool plain = !a->residual && !a->aux_bf16 && !a->out_pre_bf16;
    if (a->atomic_f32 && none && plain && !a->bias && !a->out_bf16) return launch_gemm<256, E_ATOMIC>(a, st);
    if (a->atomic_f32) return launch_gemm<256, E_GENERIC>(a, st);
    if (a->colsum != nullptr && a->act != EAVIT_ACT_GELU_BWD)      // only the GELU' and generic epilogues carry column sums
      return drop ? launch_gemm<256, E_GENERIC, true>(a, st) : launch_gemm<256, E_GENERIC>(a, st);
    if (a->act == EAVIT_ACT_GELU && a->bias && a->out_pre_bf16 && a->out_bf16 && !a->residual && !a->out_f32)
      return drop ? launch_gemm<256, E_GELU_FWD, true>(a, st) : launch_gemm<256, E_GELU_FWD>(a, st);
    if (a->act == EAVIT_ACT_GELU_BWD && a->out_bf16 && !a->bias && !a->residual && !a->out_f32 && !a->out_pre_bf16)
      return drop ? launch_gemm<256, E_GELU_BWD, true>(a, st) : launch_gemm<256, E_GELU_BWD>(a, st);
    if (none && a->bias && a->residual && a->out_f32 && !a->out_bf16 && !a->out_pre_bf16 && !a->aux_bf16)
      return drop ? launch_gemm<256, E_RESID, true>(a, st) : launch_gemm<256, E_RESID>(a, st);
    if (none && plain && !drop) return launch_gemm<256, E_STORE>(a, st);
    return drop ? launch_gemm<256, E_GENERIC, true>(a, st) : launch_gemm<256, E_GENERIC>(a, st);
  }
  if (a->N > 64) return drop ? launch_gemm<128, E_GENERIC, true>(a, st) : launch_gemm<128, E_GENERIC>(a, st);
  return drop ? launch_gemm<64, E_GENERIC, true>(a, st) : launch_gemm<64, E_GENERIC>(a, st);
}
