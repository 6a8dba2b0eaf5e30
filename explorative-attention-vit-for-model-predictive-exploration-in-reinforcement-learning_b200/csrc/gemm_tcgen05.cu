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
  int cs_smem;        // column sums requested: per-warp private partial sums in shared memory
  // E_RESID_LN: LayerNorm of the freshly formed residual row (N == BN: one tile holds whole rows) -> out_bf16, mean, rstd
  const float* ln_gamma;
  const float* ln_beta;
  float* ln_mean;
  float* ln_rstd;
  float ln_eps;
};

// Fused column sums (bias gradients): the first version issued one global red.add per (warp, chunk, column) -- 6 288 adds
// onto each of the 1 024 addresses of the MLP bias gradient per launch, which serialise in L2 and cost 170 us of the 345 us
// GELU' GEMM (r2 experiment: the same epilogue without them ran in 186 us); shared-memory float atomics (a CAS loop under
// contention from the four lane quadrants) were no better.  Now every epilogue warp keeps PRIVATE partial sums in shared
// memory -- slot (warp, chunk-of-this-warp, column-in-chunk), plain read-modify-write by the one lane that owns the column --
// and flushes them with global adds only when its CTA moves to another column block or ends.  With 148 persistent CTAs and
// 1, 2 or 4 column blocks a CTA never changes its column block: 148 adds per address and launch.
constexpr int SMEM_LIMIT = 232448;                // 227 KB per CTA

template <int BN>
struct GemmCfg {
  static constexpr int B_STAGE_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int STAGES = (BN == 256) ? 3 : (BN == 128 ? 5 : 6);
  static constexpr int TMEM_COLS = 2 * BN;    // power of two >= 32 for BN in {64,128,256}
  static constexpr int EPI_BYTES = MAX_EPI_WARPS * EPI_TILE_FLOATS * 4;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;

};

// shared memory of one CTA: TMA stages | EW transpose tiles | barrier block | EW x CS_SLOTS x 32 private column sums
// LEAN epilogues (E_*_TMA) work in the accumulator's native layout and move their global operands / results as
// 128B-swizzled [32 rows x 64 columns] bf16 boxes through the TMA unit: WBUF bytes of staging per warp, no column-sum
// slots, no LayerNorm exchange -- the shared memory they free goes to operand stages.
template <int BN, int EW, bool LEAN = false, bool WS = false, int WBUF = EPI_TILE_FLOATS * 4, int XB = 0, bool PAIR = false>
struct GemmSmem {
  // PAIR (split-K weight gradients): a work item is TWO 128-row tiles of one column block -- both TMEM accumulators -- so a
  // stage holds 256 x 64 of A next to the shared 256 x 64 of B: the B operand is read once per 256 output rows.
  // WS (weight-stationary, K <= 4 k-blocks, every CTA keeps one column block): the B operand of all k-blocks is loaded once
  // into a resident region in front of an A-only ring -- the K = 256 GEMMs otherwise re-read 128 KB of weights from L2 for
  // every 128 x 256 output tile (604 MB of L2 -> SM traffic per QKV launch next to 103 MB of activations).
  static constexpr int WS_KB = 4;
  static constexpr int B_RES_BYTES = WS ? WS_KB * GemmCfg<BN>::B_STAGE_BYTES : 0;
  static constexpr int RING_STAGE_BYTES = WS ? A_STAGE_BYTES : GemmCfg<BN>::STAGE_BYTES + (PAIR ? A_STAGE_BYTES : 0);
  static constexpr int BAR_BYTES = 384;
  static constexpr int EPI_BYTES = EW * WBUF;
  static constexpr int CS_SLOTS = (BN / 32 + EW / 4 - 1) / (EW / 4);            // 32-column chunks one warp handles per tile
  static constexpr int CS_BYTES = LEAN ? XB : EW * CS_SLOTS * 32 * 4;   // LEAN: XB bytes of per-warp scratch (bias slices)
  static constexpr int LNX_BYTES = LEAN ? 0 : 2 * (EW / 4) * 128 * 8;
  static constexpr int FIT = (SMEM_LIMIT - 1024 - BAR_BYTES - B_RES_BYTES - EPI_BYTES - CS_BYTES - LNX_BYTES) / RING_STAGE_BYTES;
  static constexpr int STAGES = (LEAN || WS) ? (FIT < 4 ? FIT : 4) : GemmCfg<BN>::STAGES;
  static constexpr int OFF_RING = B_RES_BYTES;
  static constexpr int OFF_EPI = OFF_RING + STAGES * RING_STAGE_BYTES;
  static constexpr int OFF_BAR = OFF_EPI + EPI_BYTES;
  static constexpr int OFF_CS = OFF_BAR + BAR_BYTES;
  static constexpr int OFF_LNX = OFF_CS + CS_BYTES;                            // float2 [2 buffers][warps per quadrant][128 rows] row partial sums
  static constexpr int TOTAL = OFF_LNX + LNX_BYTES + 1024 /*align slack*/;
  static_assert(STAGES >= 2, "operand ring too short");
  static_assert(TOTAL <= SMEM_LIMIT, "over the 227 KB shared-memory limit of a CTA");
  static_assert((2 * STAGES + 6 + 2 * EW) * 8 <= BAR_BYTES, "barrier block");
};

__device__ __forceinline__ float apply_act(float v, int act, float a) {
  switch (act) {
    case EAVIT_ACT_GELU: return gelu_erf(v);
    case EAVIT_ACT_GELU_BWD: return v * gelu_erf_grad(a);
    case EAVIT_ACT_MUL_AUX: return v * a;
    case EAVIT_ACT_LRELU: return v > 0.f ? v : 0.01f * v;
    case EAVIT_ACT_LRELU_BWD: return a > 0.f ? v : 0.01f * v;
    case EAVIT_ACT_RELU: return fmaxf(v, 0.f);
    case EAVIT_ACT_RELU_BWD: return a > 0.f ? v : 0.f;
    default: return v;
  }
}

// Epilogue specialisations (compile-time): the epilogue is instruction-issue bound (ncu: 45 % issue-active from only 8
// warps), so the common fused forms drop every per-element runtime branch of the generic path.
enum { E_GENERIC = 0, E_STORE = 1, E_GELU_FWD = 2, E_GELU_BWD = 3, E_RESID = 4, E_ATOMIC = 5, E_GELU_FWD_D = 6, E_MUL_AUX = 7, E_RESID_LN = 8, E_STORE_TMA = 9, E_MUL_AUX_TMA = 10, E_GELU_FWD_D_TMA = 11, E_RESID_TMA = 12, E_RESID_LN_TMA = 13 };
// E_RESID_LN (north_star clause 3: LayerNorm fused into the GEMM epilogue): x' = x + dropout(A W^T + b) as E_RESID, and --
// because N == BN == 256 == the model width, so the two warps of a lane quadrant hold whole rows between them -- the
// LayerNorm that consumes x' (vit.py:28 / :47 PreNorm) in the same epilogue: pass 1 stores x' and reduces sum / sum of squares
// per row (8-lane shuffles + one exchange between the two warps), the accumulator is released, pass 2 re-reads the rows this
// lane just wrote (L2-resident) and writes the normalised bf16 operand of the next GEMM plus (mean, rstd) for the backward.
// Deletes the standalone layernorm_fwd launch: 206 MB of fp32 re-read from HBM per LayerNorm.
// E_GELU_FWD_D / E_MUL_AUX (EAVIT_ACT_GELU_SAVE_GRAD / EAVIT_ACT_MUL_AUX): the forward stores gelu'(v) -- evaluated from the
// fp32 pre-activation with the exp and tail it computes anyway -- where E_GELU_FWD stores v, and the backward epilogue is
// one multiply instead of ~19 instructions and 2 MUFU per element.  These epilogues are bound by the serial instruction
// stream of each warp (r2 experiment, [201216,1024] x K=256: GELU' 325 us, multiply-only 252 us, store-only 186 us).
// The GELU epilogues are issue/latency-bound (16 instructions + 2 MUFU per element): 16 warps.  The store / residual /
// atomic epilogues are HBM-bound and want registers for loads in flight instead: 8 warps.
template <int EPI> struct EpiTraits {
  static constexpr bool RES = (EPI == E_RESID_TMA || EPI == E_RESID_LN_TMA);
  static constexpr bool LEAN = (EPI == E_STORE_TMA || EPI == E_MUL_AUX_TMA || EPI == E_GELU_FWD_D_TMA || RES);
  // per-warp scratch behind the barrier block: GELU: 64 bias values per warp (16 warps); residual (+ LayerNorm): 128 bias
  // values per warp (8 warps) | gamma, beta [2][256] | row-statistics exchange float2 [2][2][128]
  static constexpr int XB = (EPI == E_GELU_FWD_D_TMA) ? 16 * 256 : (RES ? 8 * 512 + 2048 + 4096 : 0);
  static constexpr int WBUF = (EPI == E_MUL_AUX_TMA || RES) ? 2 * EPI_TILE_FLOATS * 4 : EPI_TILE_FLOATS * 4;   // aux / residual box double-buffered
};
template <int EPI> struct EpiWarps {
  // measured at [201216, 1024] x K = 256 (r2): multiply-by-aux 250 us with 12 warps, 229 us with 16; residual (+ LayerNorm) epilogues
  // are best with 8 (142 / 183 us vs 145 / 232 us with 16: their prefetched fp32 rows need the registers)
  static constexpr int N = (EPI == E_GELU_FWD || EPI == E_GELU_FWD_D || EPI == E_STORE || EPI == E_MUL_AUX || EPI == E_GELU_FWD_D_TMA) ? 16 : (EPI == E_GELU_BWD ? 12 : 8);
};

template <int BN, int EPI, bool DROP, int EW = EpiWarps<EPI>::N, bool WS = false, bool PAIR = false>      // DROP: dropout mask in the epilogue (compile-time: the branch costs the fused epilogues 6-17 %)
__global__ void __launch_bounds__((CTRL_WARPS + EW) * 32, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmD,
                         const __grid_constant__ CUtensorMap tmE, const GemmKernelParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int EPI_WARPS = EW;
  using Sm = GemmSmem<BN, EPI_WARPS, EpiTraits<EPI>::LEAN || PAIR, WS, EpiTraits<EPI>::WBUF, EpiTraits<EPI>::XB, PAIR>;
  static_assert(!PAIR || EPI == E_ATOMIC, "paired tiles: split-K red.add epilogue only");
  constexpr int STAGES = Sm::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic on the __shared__ array keeps the address space: LDS / STS, not generic LD / ST
  float* epi_smem = reinterpret_cast<float*>(smem + Sm::OFF_EPI);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Sm::OFF_BAR);
  uint64_t* full_bar = bars;                      // [STAGES]   TMA -> MMA
  uint64_t* empty_bar = bars + STAGES;            // [STAGES]   MMA -> TMA
  uint64_t* acc_full = bars + 2 * STAGES;         // [2]        MMA -> epilogue
  uint64_t* acc_empty = bars + 2 * STAGES + 2;    // [2]        epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  uint64_t* b_full = bars + 2 * STAGES + 5;       // WS: resident B operand landed
  uint64_t* aux_full = bars + 2 * STAGES + 6;     // [EW][2]  E_MUL_AUX_TMA: aux box landed in the warp's buffer
  float* s_cs = reinterpret_cast<float*>(smem + Sm::OFF_CS);                     // behind the barrier block
  if (p.cs_smem)
    for (int i = threadIdx.x; i < Sm::CS_BYTES / 4; i += blockDim.x) s_cs[i] = 0.f;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = p.m_tiles * p.n_tiles * p.splits;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmA);
    tc::prefetch_tmap(&tmB);
    if constexpr (EpiTraits<EPI>::LEAN) tc::prefetch_tmap(&tmC);
    if constexpr (EPI == E_MUL_AUX_TMA || EPI == E_GELU_FWD_D_TMA || EpiTraits<EPI>::RES) tc::prefetch_tmap(&tmD);
    if constexpr (EPI == E_RESID_LN_TMA) tc::prefetch_tmap(&tmE);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) { tc::mbar_init(&full_bar[i], 1); tc::mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&acc_full[i], 1); tc::mbar_init(&acc_empty[i], EPI_WARPS); }
    tc::mbar_init(b_full, 1);
    if constexpr (EPI == E_MUL_AUX_TMA || EpiTraits<EPI>::RES) for (int i = 0; i < 2 * EPI_WARPS; ++i) tc::mbar_init(&aux_full[i], 1);
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
      auto load_b = [&](uint8_t* sb, uint64_t* bar, int kb, int n_blk) {
        if (!p.b_mn) {
          tc::tma_load_2d(sb, &tmB, bar, kb * BK, n_blk * BN);
        } else {
#pragma unroll
          for (int i = 0; i < BN / 64; ++i)
            tc::tma_load_2d(sb + i * 8192, &tmB, bar, n_blk * BN + i * 64, kb * BK);
        }
      };
      if constexpr (WS) {                        // gridDim.x % n_tiles == 0: this CTA's column block never changes
        tc::mbar_expect_tx(b_full, (uint32_t)(p.kb_total * Cfg::B_STAGE_BYTES));
        for (int kb = 0; kb < p.kb_total; ++kb) load_b(smem + kb * Cfg::B_STAGE_BYTES, b_full, kb, (int)(blockIdx.x % p.n_tiles));
      }
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n_blk = tile % p.n_tiles, m_blk = (tile / p.n_tiles) % p.m_tiles, split = tile / (p.n_tiles * p.m_tiles);
        const int kb0 = split * p.kb_per_split, kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          tc::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + Sm::OFF_RING + stage * Sm::RING_STAGE_BYTES;
          uint8_t* sb = sa + (PAIR ? 2 : 1) * A_STAGE_BYTES;
          tc::mbar_expect_tx(&full_bar[stage], Sm::RING_STAGE_BYTES);
          constexpr int AH = PAIR ? 2 : 1;                     // 128-row halves of A per stage
          if (!p.a_mn) {
#pragma unroll
            for (int h = 0; h < AH; ++h)
              tc::tma_load_2d(sa + h * A_STAGE_BYTES, &tmA, &full_bar[stage], kb * BK, (m_blk * AH + h) * BM);
          } else {
#pragma unroll
            for (int i = 0; i < AH * BM / 64; ++i)
              tc::tma_load_2d(sa + i * 8192, &tmA, &full_bar[stage], m_blk * AH * BM + i * 64, kb * BK);
          }
          if constexpr (!WS) load_b(sb, &full_bar[stage], kb, n_blk);
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
      if constexpr (WS) { tc::mbar_wait(b_full, 0); tc::fence_after_sync(); }
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int split = tile / (p.n_tiles * p.m_tiles);
        const int kb0 = split * p.kb_per_split, kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        tc::mbar_wait(&acc_empty[acc], acc_phase ^ 1);
        tc::fence_after_sync();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          tc::mbar_wait(&full_bar[stage], phase);
          tc::fence_after_sync();
          const uint32_t sa = tc::smem_u32(smem + Sm::OFF_RING + stage * Sm::RING_STAGE_BYTES);
          const uint32_t sb = WS ? tc::smem_u32(smem + kb * Cfg::B_STAGE_BYTES) : sa + (PAIR ? 2 : 1) * A_STAGE_BYTES;
          // K-major:  rows of 128 B, 8-row groups 1024 B apart (SBO); K advance = 32 B inside the swizzle atom.
          // MN-major: 64-element (128 B) MN chunks, 8 k-rows per 1024 B (SBO), MN blocks 8192 B apart (LBO);
          //           K advance = 16 k-rows = 2048 B.
          const uint64_t a_desc = p.a_mn ? tc::make_sdesc_sw128(sa, 8192, 1024) : tc::make_sdesc_sw128(sa, 16, 1024);
          const uint64_t b_desc = p.b_mn ? tc::make_sdesc_sw128(sb, 8192, 1024) : tc::make_sdesc_sw128(sb, 16, 1024);
          const uint32_t a_step = p.a_mn ? (2048 >> 4) : (32 >> 4);
          const uint32_t b_step = p.b_mn ? (2048 >> 4) : (32 >> 4);
#pragma unroll
          for (int h = 0; h < (PAIR ? 2 : 1); ++h) {   // PAIR: rows 0..127 -> accumulator 0, rows 128..255 -> accumulator 1
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              tc::mma_bf16_ss(d_tmem + h * BN, a_desc + (uint64_t)(h * (A_STAGE_BYTES >> 4)) + (uint64_t)(k * a_step),
                              b_desc + (uint64_t)(k * b_step), idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            }
          }
          tc::mma_commit(&empty_bar[stage]);          // frees the smem slot when these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tc::mma_commit(&acc_full[acc]);               // accumulator(s) complete -> epilogue
        if constexpr (PAIR) acc_phase ^= 1; else if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if constexpr (EPI == E_STORE_TMA) {
    // ===================================== epilogue: bf16 store through TMA =====================================
    // The accumulator is converted in its native layout (lane = row, registers = columns) and written as a 128B-swizzled
    // [32 rows x 64 columns] bf16 box into this warp's 4 KB staging buffer; one lane hands the box to the TMA unit
    // (cp.async.bulk.tensor store: address generation, full-line writes and the M / N tail clipping are the hardware's).
    // ~40 instructions per 32 x 64 box instead of ~640 on the transpose path (per-row address arithmetic, LDS, STG).
    static_assert(BN % 64 == 0 && EPI_TILE_FLOATS * 4 == 32 * 128, "one 32 x 64 bf16 box per staging buffer");
    const int e = warp - CTRL_WARPS;
    const int q = warp & 3;                           // TMEM lane quadrant
    const int par = e >> 2;
    constexpr int WQ = EPI_WARPS / 4;
    uint8_t* stg = reinterpret_cast<uint8_t*>(epi_smem) + e * (EPI_TILE_FLOATS * 4);
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile_i = blockIdx.x; tile_i < total_tiles; tile_i += gridDim.x) {
      const int n_blk = tile_i % p.n_tiles, m_blk = (tile_i / p.n_tiles) % p.m_tiles;
      tc::mbar_wait(&acc_full[acc], acc_phase);
      tc::fence_after_sync();
      const int row0 = m_blk * BM + q * 32;
#pragma unroll 1
      for (int g = par; g < BN / 64; g += WQ) {
        const int col0 = n_blk * BN + g * 64;
        if (col0 >= p.N) break;                        // warp-uniform
        if (lane == 0) tc::tma_store_wait_read<0>();   // the previous box has left the staging buffer
        __syncwarp();
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t r[32];
          tc::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + g * 64 + hh * 32), r);
          tc::tmem_ld_wait();
          if (p.bias != nullptr) {                     // warp-uniform; the same 16 bytes for every lane (L1 broadcast)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int c = col0 + hh * 32 + 4 * j;
              const float4 b4 = c < p.N ? __ldg(reinterpret_cast<const float4*>(p.bias + c)) : make_float4(0.f, 0.f, 0.f, 0.f);   // N % 8 == 0
              r[4 * j + 0] = __float_as_uint(__uint_as_float(r[4 * j + 0]) + b4.x); r[4 * j + 1] = __float_as_uint(__uint_as_float(r[4 * j + 1]) + b4.y);
              r[4 * j + 2] = __float_as_uint(__uint_as_float(r[4 * j + 2]) + b4.z); r[4 * j + 3] = __float_as_uint(__uint_as_float(r[4 * j + 3]) + b4.w);
            }
          }
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            uint4 v;
            v.x = pack_bf16x2(__uint_as_float(r[8 * c4 + 0]), __uint_as_float(r[8 * c4 + 1]));
            v.y = pack_bf16x2(__uint_as_float(r[8 * c4 + 2]), __uint_as_float(r[8 * c4 + 3]));
            v.z = pack_bf16x2(__uint_as_float(r[8 * c4 + 4]), __uint_as_float(r[8 * c4 + 5]));
            v.w = pack_bf16x2(__uint_as_float(r[8 * c4 + 6]), __uint_as_float(r[8 * c4 + 7]));
            *reinterpret_cast<uint4*>(stg + lane * 128 + (((hh * 4 + c4) ^ (lane & 7)) << 4)) = v;
          }
        }
        tc::fence_proxy_async();
        __syncwarp();
        if (lane == 0 && row0 < p.M) {
          tc::tma_store_2d(&tmC, stg, col0, row0);
          tc::tma_store_commit();
        }
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&acc_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) tc::tma_store_wait_all<0>();
  } else if constexpr (EpiTraits<EPI>::RES) {
    // ======= epilogue: x' = residual + acc + bias (fp32) [+ LayerNorm(x') -> bf16, mean, rstd]; operands through TMA =======
    // Items of a warp (lane quadrant q, column half par) in a tile: four fp32 boxes [32 rows x 32 columns] of its 128
    // columns -- the residual box is TMA-loaded one item ahead into one of the warp's two 4 KB buffers, x' is formed in
    // place (the lane owns the row: LDS 16 B, add, STS) and the buffer goes back to the TMA unit as the store source --
    // and, with LayerNorm, two bf16 boxes [32 x 64].  LayerNorm: the row sums need no shuffle (thread-per-row) and one
    // exchange with the warp that holds the other 128 columns; x' is parked in TENSOR MEMORY over the accumulator
    // (tcgen05.st) between the passes, so pass 2 reads it back with tcgen05.ld instead of from global memory.
    constexpr bool LN = EPI == E_RESID_LN_TMA;
    static_assert(EPI_WARPS == 8 && BN == 256, "two warps per lane quadrant, 128 columns each");
    const int e = warp - CTRL_WARPS;
    const int q = warp & 3;
    const int par = e >> 2;
    constexpr int BPW = 4;                                    // fp32 boxes per warp and tile
    uint8_t* buf0 = reinterpret_cast<uint8_t*>(epi_smem) + e * EpiTraits<EPI>::WBUF;
    uint64_t* my_bar = aux_full + 2 * e;
    float* my_bias = s_cs + e * 128;
    float* s_gb = s_cs + 8 * 128;                             // gamma[256], beta[256]
    float2* lnx = reinterpret_cast<float2*>(s_cs + 8 * 128 + 512);   // [2][2][128]
    if constexpr (LN) {
      const int te = threadIdx.x - CTRL_WARPS * 32;
      s_gb[te] = te < p.N ? __ldg(p.ln_gamma + te) : 0.f;                 // EPI_WARPS * 32 == 256 == BN
      s_gb[256 + te] = te < p.N ? __ldg(p.ln_beta + te) : 0.f;
      asm volatile("bar.sync 5, %0;" ::"n"(EPI_WARPS * 32) : "memory");
    }
    uint32_t ph0 = 0u, ph1 = 0u;
    int bias_nblk = -1, ln_buf = 0;
    auto box_ok = [&](int t, int k) { return t < total_tiles && k < BPW && (t % p.n_tiles) * BN + (par * BPW + k) * 32 < p.N; };
    auto res_load = [&](int t, int k, int b) {                // lane 0
      const int nb = t % p.n_tiles, mb = (t / p.n_tiles) % p.m_tiles;
      tc::mbar_expect_tx(&my_bar[b], 4096);
      tc::tma_load_2d(buf0 + b * 4096, &tmD, &my_bar[b], nb * BN + (par * BPW + k) * 32, mb * BM + q * 32);
    };
    auto next_tile = [&](int t) { t += (int)gridDim.x; while (t < total_tiles && !box_ok(t, 0)) t += (int)gridDim.x; return t; };
    int b = 0;
    {
      int t0 = blockIdx.x;
      while (t0 < total_tiles && !box_ok(t0, 0)) t0 += (int)gridDim.x;
      if (lane == 0 && t0 < total_tiles) res_load(t0, 0, 0);
    }
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile_i = blockIdx.x; tile_i < total_tiles; tile_i += gridDim.x) {
      const int n_blk = tile_i % p.n_tiles, m_blk = (tile_i / p.n_tiles) % p.m_tiles;
      const int row0 = m_blk * BM + q * 32;
      if (box_ok(tile_i, 0)) {
        if (n_blk != bias_nblk) {                              // warp-uniform: this warp's 128 bias values
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int c = n_blk * BN + par * 128 + i * 32 + lane;
            my_bias[i * 32 + lane] = c < p.N ? __ldg(p.bias + c) : 0.f;
          }
          bias_nblk = n_blk;
          __syncwarp();
        }
        tc::mbar_wait(&acc_full[acc], acc_phase);
        tc::fence_after_sync();
        uint32_t rk = 0u;
        if constexpr (DROP) rk = drop_row_key(p.drop, (uint32_t)(row0 + lane));
        float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
        for (int k = 0; k < BPW; ++k) {
          if (!box_ok(tile_i, k)) break;                       // warp-uniform
          uint8_t* stg = buf0 + b * 4096;
          // residual of the next fp32 box: this tile's, or (no LayerNorm passes in between) the next tile's first
          const bool more = box_ok(tile_i, k + 1);
          if (lane == 0) {
            if (more) { tc::tma_store_wait_read<0>(); res_load(tile_i, k + 1, b ^ 1); }
            else if (!LN) { const int nt = next_tile(tile_i); if (nt < total_tiles) { tc::tma_store_wait_read<0>(); res_load(nt, 0, b ^ 1); } }
          }
          if (b == 0) { tc::mbar_wait(&my_bar[0], ph0); ph0 ^= 1u; } else { tc::mbar_wait(&my_bar[1], ph1); ph1 ^= 1u; }
          uint32_t r[32];
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + (par * BPW + k) * 32);
          tc::tmem_ld_32x32(taddr, r);
          tc::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4* slot = reinterpret_cast<float4*>(stg + lane * 128 + ((j ^ (lane & 7)) << 4));
            const float4 rv = *slot;
            const float4 bv = *reinterpret_cast<const float4*>(my_bias + k * 32 + j * 4);
            float4 v;
            if constexpr (DROP) {                              // nn.Dropout on the Linear output, before the residual add
              float dm[4];
              drop4(p.drop, rk, (uint32_t)(n_blk * BN + (par * BPW + k) * 32 + 4 * j), dm);
              v.x = fmaf(__uint_as_float(r[4 * j + 0]) + bv.x, dm[0], rv.x); v.y = fmaf(__uint_as_float(r[4 * j + 1]) + bv.y, dm[1], rv.y);
              v.z = fmaf(__uint_as_float(r[4 * j + 2]) + bv.z, dm[2], rv.z); v.w = fmaf(__uint_as_float(r[4 * j + 3]) + bv.w, dm[3], rv.w);
            } else {
              v.x = __uint_as_float(r[4 * j + 0]) + bv.x + rv.x; v.y = __uint_as_float(r[4 * j + 1]) + bv.y + rv.y;
              v.z = __uint_as_float(r[4 * j + 2]) + bv.z + rv.z; v.w = __uint_as_float(r[4 * j + 3]) + bv.w + rv.w;
            }
            *slot = v;
            if constexpr (LN) {
              s1 += (v.x + v.y) + (v.z + v.w);
              s2 += fmaf(v.x, v.x, v.y * v.y) + fmaf(v.z, v.z, v.w * v.w);
              r[4 * j + 0] = __float_as_uint(v.x); r[4 * j + 1] = __float_as_uint(v.y);
              r[4 * j + 2] = __float_as_uint(v.z); r[4 * j + 3] = __float_as_uint(v.w);
            }
          }
          if constexpr (LN) tc::tmem_st_32x32(taddr, r);       // park x' over the accumulator for pass 2
          tc::fence_proxy_async();
          __syncwarp();
          if (lane == 0 && row0 < p.M) {
            tc::tma_store_2d(&tmC, stg, n_blk * BN + (par * BPW + k) * 32, row0);
            tc::tma_store_commit();
          }
          b ^= 1;
        }
        if constexpr (LN) {
          // ---- row statistics: this lane's 128 columns + the other warp's
          float2* mine = lnx + (ln_buf * 2 + par) * 128 + q * 32;
          mine[lane] = make_float2(s1, s2);
          tc::tmem_st_wait();
          asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
          const float2 t0 = lnx[(ln_buf * 2 + 0) * 128 + q * 32 + lane], t1 = lnx[(ln_buf * 2 + 1) * 128 + q * 32 + lane];
          ln_buf ^= 1;
          const float inv_n = 1.0f / (float)BN;
          const float mean = (t0.x + t1.x) * inv_n;
          const float var = fmaxf((t0.y + t1.y) * inv_n - mean * mean, 0.f);
          const float rstd = rsqrtf(var + p.ln_eps);
          if (par == 0 && p.ln_mean != nullptr && row0 + lane < p.M) { p.ln_mean[row0 + lane] = mean; p.ln_rstd[row0 + lane] = rstd; }
          const float sa = rstd, sb = -mean * rstd;
          // ---- pass 2: y = (x' - mean) * rstd * gamma + beta -> bf16, two [32 x 64] boxes
#pragma unroll 1
          for (int kk = 0; kk < 2; ++kk) {
            uint8_t* stg = buf0 + b * 4096;
            if (lane == 0) {
              tc::tma_store_wait_read<0>();                    // both buffers are free again
              if (kk == 1) { const int nt = next_tile(tile_i); if (nt < total_tiles) res_load(nt, 0, b ^ 1); }
            }
            __syncwarp();
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              uint32_t r[32];
              const int cc = par * 128 + kk * 64 + hh * 32;
              tc::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + cc), r);
              tc::tmem_ld_wait();
#pragma unroll
              for (int c4 = 0; c4 < 4; ++c4) {
                float y[8];
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2) {
                  const float4 g4 = *reinterpret_cast<const float4*>(s_gb + cc + c4 * 8 + h2 * 4);
                  const float4 be4 = *reinterpret_cast<const float4*>(s_gb + 256 + cc + c4 * 8 + h2 * 4);
                  y[4 * h2 + 0] = fmaf(fmaf(__uint_as_float(r[8 * c4 + 4 * h2 + 0]), sa, sb), g4.x, be4.x);
                  y[4 * h2 + 1] = fmaf(fmaf(__uint_as_float(r[8 * c4 + 4 * h2 + 1]), sa, sb), g4.y, be4.y);
                  y[4 * h2 + 2] = fmaf(fmaf(__uint_as_float(r[8 * c4 + 4 * h2 + 2]), sa, sb), g4.z, be4.z);
                  y[4 * h2 + 3] = fmaf(fmaf(__uint_as_float(r[8 * c4 + 4 * h2 + 3]), sa, sb), g4.w, be4.w);
                }
                *reinterpret_cast<uint4*>(stg + lane * 128 + (((hh * 4 + c4) ^ (lane & 7)) << 4)) =
                    make_uint4(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]), pack_bf16x2(y[4], y[5]), pack_bf16x2(y[6], y[7]));
              }
            }
            tc::fence_proxy_async();
            __syncwarp();
            if (lane == 0 && row0 < p.M) {
              tc::tma_store_2d(&tmE, stg, par * 128 + kk * 64, row0);
              tc::tma_store_commit();
            }
            b ^= 1;
          }
        }
      } else {
        tc::mbar_wait(&acc_full[acc], acc_phase);
        tc::fence_after_sync();
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&acc_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) tc::tma_store_wait_all<0>();
  } else if constexpr (EPI == E_GELU_FWD_D_TMA) {
    // ============ epilogue: h = gelu(acc + bias), g = gelu'(acc + bias), two bf16 outputs through TMA stores ============
    // Native layout: a lane owns a row, the bias of the warp's 64 columns sits in a warp-private shared-memory slice
    // (broadcast LDS.128).  The two boxes go through the warp's one staging buffer back to back: h is written and handed
    // to the TMA unit while g waits in 32 packed registers, then g follows as soon as the unit has read h.
    const int e = warp - CTRL_WARPS;
    const int q = warp & 3;
    const int par = e >> 2;
    constexpr int WQ = EPI_WARPS / 4;
    uint8_t* stg = reinterpret_cast<uint8_t*>(epi_smem) + e * EpiTraits<EPI>::WBUF;
    float* my_bias = s_cs + e * 64;
    int bias_col0 = -1;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile_i = blockIdx.x; tile_i < total_tiles; tile_i += gridDim.x) {
      const int n_blk = tile_i % p.n_tiles, m_blk = (tile_i / p.n_tiles) % p.m_tiles;
      const int row0 = m_blk * BM + q * 32;
      bool waited = false;
#pragma unroll 1
      for (int g = par; g < BN / 64; g += WQ) {
        const int col0 = n_blk * BN + g * 64;
        if (col0 >= p.N) break;                        // warp-uniform
        if (col0 != bias_col0) {                       // warp-uniform: a new 64-column slice of the bias
          __syncwarp();
          my_bias[lane] = col0 + lane < p.N ? __ldg(p.bias + col0 + lane) : 0.f;
          my_bias[lane + 32] = col0 + 32 + lane < p.N ? __ldg(p.bias + col0 + 32 + lane) : 0.f;
          bias_col0 = col0;
          __syncwarp();
        }
        if (!waited) { tc::mbar_wait(&acc_full[acc], acc_phase); tc::fence_after_sync(); waited = true; }
        uint32_t gp[32];                               // gelu' of the box, packed bf16 pairs
        uint32_t rk = 0u;
        if constexpr (DROP) rk = drop_row_key(p.drop, (uint32_t)(row0 + lane));
        if (lane == 0) tc::tma_store_wait_read<0>();   // the previous box (g of the last item) has left the staging buffer
        __syncwarp();
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t r[32];
          tc::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + g * 64 + hh * 32), r);
          tc::tmem_ld_wait();
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            const float4 b0 = *reinterpret_cast<const float4*>(my_bias + hh * 32 + c4 * 8);
            const float4 b1 = *reinterpret_cast<const float4*>(my_bias + hh * 32 + c4 * 8 + 4);
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            float hv[8], gv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) hv[i] = gelu_erf_and_grad(__uint_as_float(r[8 * c4 + i]) + bb[i], gv[i]);
            if constexpr (DROP) {                              // nn.Dropout after the activation (gelu' is stored unmasked)
              float dm[8];
              drop4(p.drop, rk, (uint32_t)(col0 + hh * 32 + c4 * 8), dm);
              drop4(p.drop, rk, (uint32_t)(col0 + hh * 32 + c4 * 8 + 4), dm + 4);
#pragma unroll
              for (int i = 0; i < 8; ++i) hv[i] *= dm[i];
            }
            *reinterpret_cast<uint4*>(stg + lane * 128 + (((hh * 4 + c4) ^ (lane & 7)) << 4)) =
                make_uint4(pack_bf16x2(hv[0], hv[1]), pack_bf16x2(hv[2], hv[3]), pack_bf16x2(hv[4], hv[5]), pack_bf16x2(hv[6], hv[7]));
#pragma unroll
            for (int i = 0; i < 4; ++i) gp[hh * 16 + c4 * 4 + i] = pack_bf16x2(gv[2 * i], gv[2 * i + 1]);
          }
        }
        tc::fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          if (row0 < p.M) { tc::tma_store_2d(&tmC, stg, col0, row0); tc::tma_store_commit(); }
          tc::tma_store_wait_read<0>();
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<uint4*>(stg + lane * 128 + ((j ^ (lane & 7)) << 4)) = make_uint4(gp[4 * j], gp[4 * j + 1], gp[4 * j + 2], gp[4 * j + 3]);
        tc::fence_proxy_async();
        __syncwarp();
        if (lane == 0 && row0 < p.M) { tc::tma_store_2d(&tmD, stg, col0, row0); tc::tma_store_commit(); }
      }
      if (!waited) { tc::mbar_wait(&acc_full[acc], acc_phase); tc::fence_after_sync(); }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&acc_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) tc::tma_store_wait_all<0>();
  } else if constexpr (EPI == E_MUL_AUX_TMA) {
    // ============ epilogue: out = acc * aux (the stored gelu'), bf16, + column sums; operands through TMA ============
    // A work item is a [32 rows x 64 columns] box.  The aux box is TMA-loaded into one of the warp's two 4 KB buffers one
    // item ahead; the product is formed IN PLACE (LDS 16 B of aux, multiply by the accumulator row this lane owns, STS to
    // the same address) and the buffer is handed back to the TMA unit as the store source.  The bias gradient (column sums
    // of the stored values) is read back column-wise from the finished box -- lane l owns columns 2l, 2l+1, 32 conflict-
    // free LDS.32 -- and kept in registers until the CTA leaves its column block.
    const int e = warp - CTRL_WARPS;
    const int q = warp & 3;
    const int par = e >> 2;
    constexpr int WQ = EPI_WARPS / 4;
    constexpr int NG = BN / 64;                               // column groups per tile
    constexpr int GPW = (NG + WQ - 1) / WQ;                   // groups per warp and tile
    uint8_t* buf0 = reinterpret_cast<uint8_t*>(epi_smem) + e * EpiTraits<EPI>::WBUF;
    uint64_t* my_bar = aux_full + 2 * e;
    uint32_t ph[2] = {0u, 0u};
    float cs[GPW][2];
#pragma unroll
    for (int k = 0; k < GPW; ++k) { cs[k][0] = 0.f; cs[k][1] = 0.f; }
    int cs_nblk = -1;
    auto cs_flush = [&]() {
      if (cs_nblk >= 0 && p.colsum != nullptr) {
#pragma unroll
        for (int k = 0; k < GPW; ++k) {
          const int col = cs_nblk * BN + (par + k * WQ) * 64 + 2 * lane;
          if (par + k * WQ < NG && col < p.N) { atomicAdd(p.colsum + col, cs[k][0]); atomicAdd(p.colsum + col + 1, cs[k][1]); }
          cs[k][0] = 0.f; cs[k][1] = 0.f;
        }
      }
    };
    // item iterator: (tile, k) with group g = par + k * WQ; items whose first column is past N do not exist
    auto item_ok = [&](int t, int k) { return t < total_tiles && k < GPW && par + k * WQ < NG && (t % p.n_tiles) * BN + (par + k * WQ) * 64 < p.N; };
    auto aux_load = [&](int t, int k, int b) {                // lane 0
      const int nb = t % p.n_tiles, mb = (t / p.n_tiles) % p.m_tiles;
      tc::mbar_expect_tx(&my_bar[b], 4096);
      tc::tma_load_2d(buf0 + b * 4096, &tmD, &my_bar[b], nb * BN + (par + k * WQ) * 64, mb * BM + q * 32);
    };
    // prefetch cursor: always the item after the one being processed (tiles in which this warp has no item are skipped)
    int pt = blockIdx.x, pk = 0;
    auto settle = [&]() { while (pt < total_tiles && !item_ok(pt, pk)) { pt += (int)gridDim.x; pk = 0; } };
    settle();
    int b = 0;
    if (lane == 0 && pt < total_tiles) aux_load(pt, pk, 0);
    ++pk; settle();
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile_i = blockIdx.x; tile_i < total_tiles; tile_i += gridDim.x) {
      const int n_blk = tile_i % p.n_tiles, m_blk = (tile_i / p.n_tiles) % p.m_tiles;
      if (n_blk != cs_nblk) { cs_flush(); cs_nblk = n_blk; }
      tc::mbar_wait(&acc_full[acc], acc_phase);
      tc::fence_after_sync();
      const int row0 = m_blk * BM + q * 32;
#pragma unroll
      for (int k = 0; k < GPW; ++k) {
        if (!item_ok(tile_i, k)) break;                        // warp-uniform
        const int g = par + k * WQ;
        uint8_t* stg = buf0 + b * 4096;
        if (pt < total_tiles) {
          if (lane == 0) {
            tc::tma_store_wait_read<0>();                      // the store that last read the other buffer (item - 1) is done with it
            aux_load(pt, pk, b ^ 1);
          }
          ++pk; settle();
        }
        tc::mbar_wait(&my_bar[b], ph[b]);
        ph[b] ^= 1u;
        uint32_t rk = 0u;
        if constexpr (DROP) rk = drop_row_key(p.drop, (uint32_t)(row0 + lane));
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t r[32];
          tc::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + g * 64 + hh * 32), r);
          tc::tmem_ld_wait();
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            uint4* slot = reinterpret_cast<uint4*>(stg + lane * 128 + (((hh * 4 + c4) ^ (lane & 7)) << 4));
            const uint4 a = *slot;
            if constexpr (DROP) {                              // the forward's dropout mask on the MLP hidden, regenerated
              float dm[8];
              drop4(p.drop, rk, (uint32_t)(n_blk * BN + g * 64 + hh * 32 + c4 * 8), dm);
              drop4(p.drop, rk, (uint32_t)(n_blk * BN + g * 64 + hh * 32 + c4 * 8 + 4), dm + 4);
#pragma unroll
              for (int i = 0; i < 8; ++i) r[8 * c4 + i] = __float_as_uint(__uint_as_float(r[8 * c4 + i]) * dm[i]);
            }
            uint4 v;
            v.x = pack_bf16x2(__uint_as_float(r[8 * c4 + 0]) * __uint_as_float(a.x << 16), __uint_as_float(r[8 * c4 + 1]) * __uint_as_float(a.x & 0xffff0000u));
            v.y = pack_bf16x2(__uint_as_float(r[8 * c4 + 2]) * __uint_as_float(a.y << 16), __uint_as_float(r[8 * c4 + 3]) * __uint_as_float(a.y & 0xffff0000u));
            v.z = pack_bf16x2(__uint_as_float(r[8 * c4 + 4]) * __uint_as_float(a.z << 16), __uint_as_float(r[8 * c4 + 5]) * __uint_as_float(a.z & 0xffff0000u));
            v.w = pack_bf16x2(__uint_as_float(r[8 * c4 + 6]) * __uint_as_float(a.w << 16), __uint_as_float(r[8 * c4 + 7]) * __uint_as_float(a.w & 0xffff0000u));
            *slot = v;
          }
        }
        tc::fence_proxy_async();
        __syncwarp();
        if (lane == 0 && row0 < p.M) {
          tc::tma_store_2d(&tmC, stg, n_blk * BN + g * 64, row0);
          tc::tma_store_commit();
        }
        if (p.colsum != nullptr) {                             // rows past M were zero-filled by the aux load: they add 0
          float s0 = 0.f, s1 = 0.f;
          const int ch = lane >> 2, w4 = (lane & 3) << 2;
#pragma unroll
          for (int rr = 0; rr < 32; ++rr) {
            const uint32_t u = *reinterpret_cast<const uint32_t*>(stg + rr * 128 + ((ch ^ (rr & 7)) << 4) + w4);
            s0 += __uint_as_float(u << 16);
            s1 += __uint_as_float(u & 0xffff0000u);
          }
          cs[k][0] += s0; cs[k][1] += s1;
        }
        __syncwarp();
        b ^= 1;
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&acc_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    cs_flush();
    if (lane == 0) tc::tma_store_wait_all<0>();
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
    const bool need_aux = GEN ? (p.act == EAVIT_ACT_GELU_BWD || p.act == EAVIT_ACT_LRELU_BWD || p.act == EAVIT_ACT_RELU_BWD || p.act == EAVIT_ACT_MUL_AUX)
                              : (EPI == E_GELU_BWD || EPI == E_MUL_AUX);
    constexpr bool LNF = EPI == E_RESID_LN;
    const bool has_bias = GEN || EPI == E_STORE ? (p.bias != nullptr) : (EPI == E_GELU_FWD || EPI == E_GELU_FWD_D || EPI == E_RESID || LNF);
    const bool has_pre = GEN ? (p.out_pre != nullptr) : (EPI == E_GELU_FWD);
    const bool has_res = GEN ? (p.residual != nullptr) : (EPI == E_RESID || LNF);
    const bool out_f32 = GEN || EPI == E_STORE ? (p.out_f32 != nullptr) : (EPI == E_RESID || EPI == E_ATOMIC || LNF);
    const bool out_b16 = GEN || EPI == E_STORE ? (p.out_bf16 != nullptr) : (EPI == E_GELU_FWD || EPI == E_GELU_FWD_D || EPI == E_GELU_BWD || EPI == E_MUL_AUX);
    const bool atomic = GEN ? (p.atomic_f32 != 0) : (EPI == E_ATOMIC);
    constexpr int CS_SLOTS = Sm::CS_SLOTS;                                         // chunks one warp handles per tile
    float* my_cs = s_cs + e * CS_SLOTS * 32;                                       // this warp's private partial sums
    int cs_nblk = -1;                                                              // column block the partial sums belong to
    auto cs_flush = [&]() {                                                        // warp-uniform
      if (cs_nblk >= 0) {
#pragma unroll 1
        for (int k = 0; k < CS_SLOTS; ++k) {
          const int col = cs_nblk * BN + (par + k * (EPI_WARPS / 4)) * 32 + lane;
          const float t = my_cs[k * 32 + lane];
          my_cs[k * 32 + lane] = 0.f;
          if (par + k * (EPI_WARPS / 4) < BN / 32 && col < p.N && t != 0.f) atomicAdd(p.colsum + col, t);
        }
      }
      __syncwarp();
    };
    float2* lnx = reinterpret_cast<float2*>(smem + Sm::OFF_LNX);
    int ln_buf = 0;
    int acc = 0; uint32_t acc_phase = 0;
    // Global operands of the epilogue (residual / aux) do not depend on the accumulator: the loads of a 32 x 32 chunk are
    // issued one chunk AHEAD -- into a second register set where the budget allows (PF2), and for every variant the first
    // chunk of a tile before the wait on its accumulator -- so an HBM round trip is hidden behind the previous chunk's
    // TMEM read / transpose / stores instead of sitting at the head of every chunk.
    constexpr bool PF2 = false;   // measured: the second register set spills at 96 registers (MUL_AUX 233 -> 299 us) and costs RESID 145 -> 153 us
    constexpr int WQ = EPI_WARPS / 4;                                              // warps per lane quadrant
    float4 res[8], resN[PF2 ? 8 : 1];
    uint2 ax[8], axN[PF2 ? 8 : 1];
    auto issue = [&](int t, int c, float4* R, uint2* A) {
      if (t >= total_tiles || c >= BN / 32) return;
      const int nb = t % p.n_tiles, mb = (t / p.n_tiles) % p.m_tiles;
      const int colL = nb * BN + c * 32 + (lane & 7) * 4;
      if (colL >= p.N) return;
      const int r0 = mb * BM + q * 32, nr = p.M - r0, rs = lane >> 3;
      const size_t o0 = (size_t)(r0 + rs) * (size_t)p.ldc + colL, os = 4 * (size_t)p.ldc;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        if (nr >= 32 || it * 4 + rs < nr) {
          if (has_res) R[it] = __ldg(reinterpret_cast<const float4*>(p.residual + o0 + it * os));
          if (need_aux) A[it] = __ldg(reinterpret_cast<const uint2*>(p.aux + o0 + it * os));
        }
      }
    };
    const bool ep_loads = has_res || need_aux;
    bool primed = false;                                  // res / ax already hold the first chunk of the upcoming tile
    for (int tile_i = blockIdx.x; tile_i < total_tiles; tile_i += gridDim.x) {
      const int n_blk = tile_i % p.n_tiles, m_blk = (tile_i / p.n_tiles) % p.m_tiles;
      if (p.cs_smem && n_blk != cs_nblk) { cs_flush(); cs_nblk = n_blk; }
      if (ep_loads && !primed) issue(tile_i, par, res, ax);
      primed = false;
      float ls1[8], ls2[8];                               // E_RESID_LN: partial row sums of rows rsub + 4 it over this warp's columns
#pragma unroll
      for (int it = 0; it < 8; ++it) { ls1[it] = 0.f; ls2[it] = 0.f; }
      tc::mbar_wait(&acc_full[acc], acc_phase);
      tc::fence_after_sync();
#pragma unroll 1
      for (int half = 0; half < (PAIR ? 2 : 1); ++half) {
      const int acc_c = PAIR ? half : acc;             // accumulator read below
      const int row0 = (PAIR ? m_blk * 2 + half : m_blk) * BM + q * 32;
      const int nrows = min(32, p.M - row0);          // may be <= 0 for the M tail
#pragma unroll 1
      for (int c = par; c < BN / 32; c += EPI_WARPS / 4) {
        const int col0 = n_blk * BN + c * 32;
        if (col0 >= p.N) break;                        // warp-uniform
        const int cchunk = lane & 7, rsub = lane >> 3;
        const int col = col0 + cchunk * 4;
        constexpr bool CS = GEN || EPI == E_GELU_BWD || EPI == E_MUL_AUX;    // epilogues that can carry a fused column sum
        const size_t off0 = (size_t)(row0 + rsub) * (size_t)p.ldc + col;      // row rr = rsub + 4 * it
        const size_t ostep = 4 * (size_t)p.ldc;
        const int cn = c + WQ;
        const bool more = cn < BN / 32 && n_blk * BN + cn * 32 < p.N;        // this warp has another chunk in this tile
        if constexpr (PF2) {
          if (ep_loads) { if (more) issue(tile_i, cn, resN, axN); else issue(tile_i + (int)gridDim.x, par, resN, axN); }
        }
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (col < p.N && has_bias) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
        // registers (one row per lane) -> swizzled smem tile -> (4 rows x 8 lanes x 16 B) per instruction; two 16-column
        // halves so that only 16 accumulator registers are live next to the prefetched residual / aux values
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t r[16];
          tc::tmem_ld_32x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc_c * BN + c * 32 + hh * 16), r);
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
            if (has_pre && !(GEN && p.act == EAVIT_ACT_GELU_SAVE_GRAD))
              *reinterpret_cast<uint2*>(p.out_pre + off) = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
            if constexpr (EPI == E_GELU_FWD) {
#pragma unroll
              for (int i = 0; i < 4; ++i) v[i] = gelu_erf(v[i]);
            } else if constexpr (EPI == E_GELU_FWD_D) {
              float g[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) v[i] = gelu_erf_and_grad(v[i], g[i]);
              *reinterpret_cast<uint2*>(p.out_pre + off) = make_uint2(pack_bf16x2(g[0], g[1]), pack_bf16x2(g[2], g[3]));
            } else if constexpr (EPI == E_MUL_AUX) {
              v[0] *= __uint_as_float(ax[it].x << 16); v[1] *= __uint_as_float(ax[it].x & 0xffff0000u);
              v[2] *= __uint_as_float(ax[it].y << 16); v[3] *= __uint_as_float(ax[it].y & 0xffff0000u);
            } else if constexpr (EPI == E_GELU_BWD) {
              v[0] *= gelu_erf_grad(__uint_as_float(ax[it].x << 16)); v[1] *= gelu_erf_grad(__uint_as_float(ax[it].x & 0xffff0000u));
              v[2] *= gelu_erf_grad(__uint_as_float(ax[it].y << 16)); v[3] *= gelu_erf_grad(__uint_as_float(ax[it].y & 0xffff0000u));
            } else if constexpr (GEN) {
              if (p.act == EAVIT_ACT_GELU_SAVE_GRAD) {
                float g[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) v[i] = gelu_erf_and_grad(v[i], g[i]);
                if (p.out_pre != nullptr)
                  *reinterpret_cast<uint2*>(p.out_pre + off) = make_uint2(pack_bf16x2(g[0], g[1]), pack_bf16x2(g[2], g[3]));
              } else if (p.act != EAVIT_ACT_NONE) {
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
            if constexpr (LNF) {
              ls1[it] += (v[0] + v[1]) + (v[2] + v[3]);
              ls2[it] += fmaf(v[0], v[0], v[1] * v[1]) + fmaf(v[2], v[2], v[3] * v[3]);
            }
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
        if (ep_loads) {
          if constexpr (PF2) {
#pragma unroll
            for (int it = 0; it < 8; ++it) { res[it] = resN[it]; ax[it] = axN[it]; }
            if (!more) primed = true;
          } else {
            if (more) issue(tile_i, cn, res, ax);
          }
        }
        if (p.colsum != nullptr) {                         // warp-uniform
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            cs[i] += __shfl_xor_sync(0xffffffffu, cs[i], 8);
            cs[i] += __shfl_xor_sync(0xffffffffu, cs[i], 16);
          }
          if (rsub == 0 && col < p.N) {                 // lanes 0..7 own columns cchunk*4 .. +3 of slot (c - par) / warps-per-quadrant
            float4* slot = reinterpret_cast<float4*>(my_cs + ((c - par) / (EPI_WARPS / 4)) * 32 + cchunk * 4);
            float4 t = *slot;
            t.x += cs[0]; t.y += cs[1]; t.z += cs[2]; t.w += cs[3];
            *slot = t;
          }
        }
        __syncwarp();
      }
      }   // half
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&acc_empty[acc]);
      if constexpr (PAIR) acc_phase ^= 1; else if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      if constexpr (LNF) {
        // ---- row statistics: 8 lanes share a row inside the warp, the two warps of the quadrant share it between them
        const int cchunk = lane & 7, rsub = lane >> 3;
        const int row0 = m_blk * BM + q * 32;
        const int nrows = min(32, p.M - row0);
#pragma unroll
        for (int it = 0; it < 8; ++it) {
#pragma unroll
          for (int o = 1; o < 8; o <<= 1) {
            ls1[it] += __shfl_xor_sync(0xffffffffu, ls1[it], o);
            ls2[it] += __shfl_xor_sync(0xffffffffu, ls2[it], o);
          }
        }
        float2* mine = lnx + (ln_buf * WQ + par) * 128 + q * 32;
        if (cchunk == 0) {
#pragma unroll
          for (int it = 0; it < 8; ++it) mine[it * 4 + rsub] = make_float2(ls1[it], ls2[it]);
        }
        asm volatile("bar.sync %0, %1;" ::"r"(1 + q), "n"(WQ * 32) : "memory");          // the warps of lane quadrant q
        float mean[8], rstd[8];
        const float inv_n = 1.0f / (float)BN;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          float2 o2 = make_float2(0.f, 0.f);
#pragma unroll
          for (int w2 = 0; w2 < WQ; ++w2) {                                 // same order in every warp: identical statistics
            const float2 t2 = lnx[(ln_buf * WQ + w2) * 128 + q * 32 + it * 4 + rsub];
            o2.x += t2.x; o2.y += t2.y;
          }
          const float m = o2.x * inv_n;
          const float var = fmaxf(o2.y * inv_n - m * m, 0.f);
          mean[it] = m;
          rstd[it] = rsqrtf(var + p.ln_eps);
        }
        ln_buf ^= 1;
        if (par == 0 && cchunk == 0 && p.ln_mean != nullptr) {
#pragma unroll
          for (int it = 0; it < 8; ++it)
            if (it * 4 + rsub < nrows) { p.ln_mean[row0 + it * 4 + rsub] = mean[it]; p.ln_rstd[row0 + it * 4 + rsub] = rstd[it]; }
        }
        // ---- pass 2: y = (x' - mean) * rstd * gamma + beta -> bf16, from the rows this lane stored in pass 1
#pragma unroll 1
        for (int c = par; c < BN / 32; c += EPI_WARPS / 4) {
          const int col = c * 32 + cchunk * 4;
          const float4 g4 = __ldg(reinterpret_cast<const float4*>(p.ln_gamma + col));
          const float4 be4 = __ldg(reinterpret_cast<const float4*>(p.ln_beta + col));
          const size_t off0 = (size_t)(row0 + rsub) * (size_t)p.ldc + col;
          const size_t ostep = 4 * (size_t)p.ldc;
          float4 xv[8];
#pragma unroll
          for (int it = 0; it < 8; ++it)
            if (it * 4 + rsub < nrows) xv[it] = __ldcg(reinterpret_cast<const float4*>(p.out_f32 + off0 + it * ostep));
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            if (it * 4 + rsub < nrows) {
              const float a = rstd[it], b = -mean[it] * rstd[it];
              const float y0 = fmaf(fmaf(xv[it].x, a, b), g4.x, be4.x), y1 = fmaf(fmaf(xv[it].y, a, b), g4.y, be4.y);
              const float y2 = fmaf(fmaf(xv[it].z, a, b), g4.z, be4.z), y3 = fmaf(fmaf(xv[it].w, a, b), g4.w, be4.w);
              *reinterpret_cast<uint2*>(p.out_bf16 + off0 + it * ostep) = make_uint2(pack_bf16x2(y0, y1), pack_bf16x2(y2, y3));
            }
          }
        }
      }
    }
    if (p.cs_smem) cs_flush();
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

// 2-D fp32 tensor map, 128-byte swizzle: dims {inner, outer}, box {32 floats = 128 B, box_outer}
static int make_tmap_f32_2d(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_bytes, uint32_t box_outer) {
  EncodeTiledFn enc = get_encode_fn();
  if (enc == nullptr) return EAVIT_ECUDA;
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) { cudaFree(nullptr); ctx_bound = true; }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_bytes};
  cuuint32_t box[2] = {32, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (fp32) failed: %d (base %p inner %llu outer %llu pitch %llu)", (int)r, base,
              (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)pitch_bytes);
    return EAVIT_ECUDA;
  }
  return EAVIT_OK;
}

template <int BN, int EPI, bool DROP = false, int EW = EpiWarps<EPI>::N, bool WS = false, bool PAIR = false>
static int launch_gemm(const eavit_gemm_args* a, cudaStream_t st) {
  using Cfg = GemmCfg<BN>;
  static bool attr_done = false;
  if (!attr_done) {
    EAVIT_CUDA(cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<BN, EPI, DROP, EW, WS, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    GemmSmem<BN, EW, EpiTraits<EPI>::LEAN || PAIR, WS, EpiTraits<EPI>::WBUF, EpiTraits<EPI>::XB, PAIR>::TOTAL));
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
  CUtensorMap tmC = tmA, tmD = tmA;          // only the TMA epilogues read them
  if (EPI == E_MUL_AUX_TMA) {
    rc = make_tmap_bf16_2d(&tmD, a->aux_bf16, (uint64_t)a->N, (uint64_t)a->M, (uint64_t)a->ldc * 2, 32);
    if (rc) return rc;
  }
  if (EPI == E_GELU_FWD_D_TMA) {
    rc = make_tmap_bf16_2d(&tmD, a->out_pre_bf16, (uint64_t)a->N, (uint64_t)a->M, (uint64_t)a->ldc * 2, 32);
    if (rc) return rc;
  }
  CUtensorMap tmE = tmA;
  if (EpiTraits<EPI>::RES) {
    rc = make_tmap_f32_2d(&tmD, a->residual, (uint64_t)a->N, (uint64_t)a->M, (uint64_t)a->ldc * 4, 32);
    if (rc) return rc;
    rc = make_tmap_f32_2d(&tmC, a->out_f32, (uint64_t)a->N, (uint64_t)a->M, (uint64_t)a->ldc * 4, 32);
    if (rc) return rc;
    if (EPI == E_RESID_LN_TMA) {
      rc = make_tmap_bf16_2d(&tmE, a->out_bf16, (uint64_t)a->N, (uint64_t)a->M, (uint64_t)a->ldc * 2, 32);
      if (rc) return rc;
    }
  } else if (EpiTraits<EPI>::LEAN) {
    rc = make_tmap_bf16_2d(&tmC, a->out_bf16, (uint64_t)a->N, (uint64_t)a->M, (uint64_t)a->ldc * 2, 32);
    if (rc) return rc;
  }

  GemmKernelParams p;
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.a_mn = a->a_mn; p.b_mn = a->b_mn;
  p.m_tiles = PAIR ? cdiv(cdiv(a->M, BM), 2) : cdiv(a->M, BM);      // PAIR: work items are pairs of 128-row tiles
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
  p.cs_smem = a->colsum != nullptr ? 1 : 0;
  p.ln_gamma = a->ln_gamma; p.ln_beta = a->ln_beta; p.ln_mean = a->ln_mean; p.ln_rstd = a->ln_rstd; p.ln_eps = a->ln_eps;
  const int total = p.m_tiles * p.n_tiles * p.splits;
  int grid = total < kNumSMs ? total : kNumSMs;
  if (WS) grid = grid / p.n_tiles * p.n_tiles;           // a CTA keeps its column block: tile % n_tiles == blockIdx.x % n_tiles
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3((CTRL_WARPS + EW) * 32);
  cfg.dynamicSmemBytes = GemmSmem<BN, EW, EpiTraits<EPI>::LEAN || PAIR, WS, EpiTraits<EPI>::WBUF, EpiTraits<EPI>::XB, PAIR>::TOTAL;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  EAVIT_CUDA(cudaLaunchKernelEx(&cfg, gemm_bf16_tcgen05_kernel<BN, EPI, DROP, EW, WS, PAIR>, tmA, tmB, tmC, tmD, tmE, p));
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
  const bool need_aux = (a->act == EAVIT_ACT_GELU_BWD || a->act == EAVIT_ACT_LRELU_BWD || a->act == EAVIT_ACT_RELU_BWD || a->act == EAVIT_ACT_MUL_AUX);
  EAVIT_CHECK_ARG(!need_aux || a->aux_bf16 != nullptr);
  EAVIT_CHECK_ARG(a->act >= EAVIT_ACT_NONE && a->act <= EAVIT_ACT_GELU_SAVE_GRAD);
  if (a->ln_gamma != nullptr) {            // fused LayerNorm of the residual row
    EAVIT_CHECK_ARG(a->N == 256 && a->ln_beta && a->bias && a->residual && a->out_f32 && a->out_bf16 && !a->out_pre_bf16 && !a->aux_bf16);
    EAVIT_CHECK_ARG(a->act == EAVIT_ACT_NONE && !a->atomic_f32 && a->split_k <= 1 && !a->colsum && (a->ln_mean == nullptr) == (a->ln_rstd == nullptr));
    EAVIT_CHECK_ARG(a->out_f32 != a->residual || true);
  }
  EAVIT_CHECK_ARG(a->bias == nullptr || (reinterpret_cast<uintptr_t>(a->bias) & 15) == 0);
  cudaStream_t st = (cudaStream_t)stream;
  // switches for A/B measurements and for falling back to the transposing epilogues (read once)
  static const bool no_tma = getenv("EAVIT_NO_TMA_STORE") != nullptr, no_ws = getenv("EAVIT_NO_WS") != nullptr,
                    no_pair = getenv("EAVIT_NO_PAIR") != nullptr;
  const bool drop = make_drop(a->drop_p, a->drop_seed).thresh != 0;
  if (a->N > 128) {
    const bool none = a->act == EAVIT_ACT_NONE;
    const bool plain = !a->residual && !a->aux_bf16 && !a->out_pre_bf16;
    if (a->atomic_f32 && none && plain && !a->bias && !a->out_bf16 && !a->colsum && !drop && a->M > BM &&
        cdiv(a->M, BM) * cdiv(a->N, 256) >= 4 && !no_pair)   // engine._split_k mirrors this rule; 2 tiles (256 x 256) are faster unpaired (62 vs 74 us)
      return launch_gemm<256, E_ATOMIC, false, 8, false, true>(a, st);
    if (a->atomic_f32 && none && plain && !a->bias && !a->out_bf16) return launch_gemm<256, E_ATOMIC>(a, st);
    if (a->atomic_f32) return launch_gemm<256, E_GENERIC>(a, st);
    if (a->colsum != nullptr && a->act != EAVIT_ACT_GELU_BWD && a->act != EAVIT_ACT_MUL_AUX)   // only the GELU' / multiply and generic epilogues carry column sums
      return drop ? launch_gemm<256, E_GENERIC, true>(a, st) : launch_gemm<256, E_GENERIC>(a, st);
    if (a->act == EAVIT_ACT_GELU && a->bias && a->out_pre_bf16 && a->out_bf16 && !a->residual && !a->out_f32)
      return drop ? launch_gemm<256, E_GELU_FWD, true>(a, st) : launch_gemm<256, E_GELU_FWD>(a, st);
    if (a->act == EAVIT_ACT_GELU_BWD && a->out_bf16 && !a->bias && !a->residual && !a->out_f32 && !a->out_pre_bf16)
      return drop ? launch_gemm<256, E_GELU_BWD, true>(a, st) : launch_gemm<256, E_GELU_BWD>(a, st);
    if (a->act == EAVIT_ACT_GELU_SAVE_GRAD && a->bias && a->out_pre_bf16 && a->out_bf16 && !a->residual && !a->out_f32 &&
        (a->ldc * 2) % 16 == 0 && (reinterpret_cast<uintptr_t>(a->out_bf16) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->out_pre_bf16) & 15) == 0 &&
        !no_tma)
      return drop ? launch_gemm<256, E_GELU_FWD_D_TMA, true>(a, st) : launch_gemm<256, E_GELU_FWD_D_TMA>(a, st);
    if (a->act == EAVIT_ACT_GELU_SAVE_GRAD && a->bias && a->out_pre_bf16 && a->out_bf16 && !a->residual && !a->out_f32)
      return drop ? launch_gemm<256, E_GELU_FWD_D, true>(a, st) : launch_gemm<256, E_GELU_FWD_D>(a, st);
    const bool tma_ok = (a->ldc * 2) % 16 == 0 && (reinterpret_cast<uintptr_t>(a->out_bf16) & 15) == 0 && !no_tma;
    if (a->act == EAVIT_ACT_MUL_AUX && a->out_bf16 && !a->bias && !a->residual && !a->out_f32 && !a->out_pre_bf16 && tma_ok &&
        (reinterpret_cast<uintptr_t>(a->aux_bf16) & 15) == 0)     // (weight-stationary B leaves 2 A stages beside the aux buffers: 181 vs 155 us)
      return drop ? launch_gemm<256, E_MUL_AUX_TMA, true, 8>(a, st) : launch_gemm<256, E_MUL_AUX_TMA, false, 8>(a, st);
    if (a->act == EAVIT_ACT_MUL_AUX && a->out_bf16 && !a->bias && !a->residual && !a->out_f32 && !a->out_pre_bf16)
      return drop ? launch_gemm<256, E_MUL_AUX, true>(a, st) : launch_gemm<256, E_MUL_AUX>(a, st);
    const bool res_tma = a->ldc % 8 == 0 && (reinterpret_cast<uintptr_t>(a->residual) & 15) == 0 &&
                         (reinterpret_cast<uintptr_t>(a->out_f32) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->out_bf16) & 15) == 0 &&
                         !no_tma;
    if (a->ln_gamma != nullptr && res_tma) return drop ? launch_gemm<256, E_RESID_LN_TMA, true, 8>(a, st) : launch_gemm<256, E_RESID_LN_TMA, false, 8>(a, st);
    if (none && a->bias && a->residual && a->out_f32 && !a->out_bf16 && !a->out_pre_bf16 && !a->aux_bf16 && !a->colsum && res_tma)
      return drop ? launch_gemm<256, E_RESID_TMA, true, 8>(a, st) : launch_gemm<256, E_RESID_TMA, false, 8>(a, st);
    if (a->ln_gamma != nullptr)            // checked above: N == 256, bias + residual + fp32 and bf16 outputs, no split-K
      return drop ? launch_gemm<256, E_RESID_LN, true>(a, st) : launch_gemm<256, E_RESID_LN>(a, st);
    if (none && a->bias && a->residual && a->out_f32 && !a->out_bf16 && !a->out_pre_bf16 && !a->aux_bf16)
      return drop ? launch_gemm<256, E_RESID, true>(a, st) : launch_gemm<256, E_RESID>(a, st);
    if (none && plain && !drop && a->out_bf16 && !a->out_f32 && !a->colsum && (a->ldc * 2) % 16 == 0 &&
        (reinterpret_cast<uintptr_t>(a->out_bf16) & 15) == 0 && !no_tma)
    {
      const int kbt = cdiv(a->K, BK), nt = cdiv(a->N, 256), tiles = cdiv(a->M, BM) * nt;
      if (kbt <= 4 && tiles >= kNumSMs && nt <= kNumSMs && !no_ws) return launch_gemm<256, E_STORE_TMA, false, 8, true>(a, st);
      return launch_gemm<256, E_STORE_TMA, false, 8>(a, st);
    }
    if (none && plain && !drop) return launch_gemm<256, E_STORE>(a, st);
    return drop ? launch_gemm<256, E_GENERIC, true>(a, st) : launch_gemm<256, E_GENERIC>(a, st);
  }
  if (a->N > 64) return drop ? launch_gemm<128, E_GENERIC, true>(a, st) : launch_gemm<128, E_GENERIC>(a, st);
  return drop ? launch_gemm<64, E_GENERIC, true>(a, st) : launch_gemm<64, E_GENERIC>(a, st);
}
