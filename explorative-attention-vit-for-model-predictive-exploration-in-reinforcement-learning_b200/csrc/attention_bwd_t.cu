// Attention backward with TRANSPOSED scores (keys on the TMEM lanes): Dh = 32 with up to 208 tokens, Dh = 64 with up to 128.
//
// attention_tc.cu's backward keeps queries on the lanes: P~ and dS must both pass through ONE shared-memory buffer (they
// are MN-major A operands of dV = P~^T dO and dK = dS^T Q), which forces  S,dP MMA -> pass 1 (P~, D) -> dV MMA -> pass 2
// (dS) -> dK,dQ MMA  with three MMA round trips and two softmax passes per (head, query tile).  With the scores transposed,
//     S^T = K Q^T,   dP^T = V dO^T              (M = 128 keys of a key tile, N = all queries of the sequence)
// P~^T and dS^T have the keys on the lanes -- exactly the layout of a TMEM-resident A operand -- so
//     dV_kt = P~^T dO   and   dK_kt = dS^T Q    read A from TENSOR MEMORY (tcgen05.mma with [tmem] A),
// written there by tcgen05.st as packed bf16 pairs over the first half of each thread's own score columns; only dS^T also
// goes to shared memory (dQ = dS K needs queries on M: the same buffer read as an MN-major A operand).  The row statistic
// D_i = sum_j P_ij dP_ij = sum_d O_id dO_id is formed from the staged O and dO tiles, so ONE pass produces P~^T and dS^T together:
//     per (head, key tile):   S^T, dP^T MMA -> one pass -> dV, dK, dQ MMAs -> drain        (two MMA round trips, one pass)
// dV / dK are complete per key tile (their K dimension is the whole sequence) and are stored straight from TMEM; dQ is
// summed over the two key tiles in registers.
#include "common.cuh"
#include "tc05.cuh"

namespace eavit {

int make_tmap_bf16_2d(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_bytes, uint32_t box_outer);

namespace bt {

// Optional phase timing (-DEAVIT_TRACE, tools/att_trace.py bwdt): clock64 deltas of block 0 / thread 0 per phase.
#ifdef EAVIT_TRACE
__device__ long long g_bt_trace[32];
#define BTR(i) do { if (threadIdx.x == 0) { const long long _t = clock64(); tr_acc[i] += _t - tr_prev; tr_prev = _t; } } while (0)
#else
#define BTR(i) do { } while (0)
#endif

constexpr int THREADS = 512;
constexpr int MAXQ = 208;                   // queries (score columns) per sequence: S^T and dP^T take 2 x 208 TMEM columns
constexpr int ROWB = 128;                   // bytes per staged row: 64 bf16 = two heads of 32
constexpr int SLAB = 128 * ROWB;            // [128 rows][64 columns] bf16, 128-byte swizzle
constexpr int BOX_ROWS = 32;
constexpr int BOX_BYTES = BOX_ROWS * ROWB;
constexpr int MAXQ64 = 128;                 // Dh = 64: one key / query tile (three 64-column accumulators beside S^T and dP^T)

__device__ __forceinline__ uint32_t sw_off(int r, int c) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
}

struct Smem {
  static constexpr int OFF_Q = 0;                               // [256][128 B]
  static constexpr int OFF_DO = OFF_Q + 256 * ROWB;             // [256][128 B]
  static constexpr int OFF_K = OFF_DO + 256 * ROWB;             // [224][128 B] (tile 1 of an MMA reads on into V / dS: finite)
  static constexpr int OFF_V = OFF_K + 224 * ROWB;
  static constexpr int OFF_O = OFF_V + 224 * ROWB;              // [224][128 B]  attention output (for D = rowsum(O * dO))
  static constexpr int OFF_DS = OFF_O + 224 * ROWB;             // dS^T: 4 slabs [128 keys][64 queries]
  static constexpr int OFF_L = OFF_DS + 4 * SLAB;               // float [2 buffers][2 heads][256]  lse * log2(e)   (+inf past the end)
  static constexpr int OFF_D = OFF_L + 2 * 2 * 256 * 4;         // float [2 heads][256]  D
  static constexpr int OFF_RK = OFF_D + 2 * 256 * 4;            // uint32 [256]  dropout row keys of the item's query rows
  static constexpr int OFF_BAR = OFF_RK + 256 * 4;
  static constexpr int TOTAL = OFF_BAR + 128 + 1024;
};

// k-steps (16 queries) of warp group g: with two query tiles the LAST group takes >= 4 so that the second half of its
// columns (free once its scores are consumed) holds the 32-column dQ accumulator of query tile 1.
__device__ __forceinline__ void ksplit(int nks, bool two_q, int* kb /*[5]*/) {
  if (two_q) {
    const int k3 = max(4, (nks + 3) >> 2), rest = nks - k3;
    kb[0] = 0; kb[1] = (rest + 2) / 3; kb[2] = kb[1] + (rest + 1) / 3; kb[3] = rest; kb[4] = nks;
  } else {
    for (int g = 0; g <= 4; ++g) kb[g] = (g * nks) >> 2;
  }
}

// DROP: nn.Dropout on the attention probabilities (vit.py:45,70): mask element = (token row t0 + q, column h*256 + key), the
// same element attention_tc.cu evaluates.  dV takes the masked P~ (the stash), dS = P (M dP - D), D = rowsum(O * dO) holds
// unchanged because O was formed from the masked probabilities.
// SEG: `pack` consecutive equal-length sequences (contiguous in the token matrix) form one work item under a block-diagonal
// mask -- a key attends only to the queries of its own `seg`-token sequence; masked probabilities are exact zeros, so every
// product is unchanged (see attention_tc.cu).  This is what makes the 50-token sequences of the HF-style ViT worth a tile.
template <int DH, bool DROP, bool SEG>
__global__ void __launch_bounds__(THREADS, 1)
attention_bwd_t_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                       const __grid_constant__ CUtensorMap tmO, const float* __restrict__ lse, const int* __restrict__ seq_start,
                       int nseq, int H, float scale, __nv_bfloat16* __restrict__ dqkv, const DropCfg drop, int pack, int seg) {
  constexpr int HPB = 64 / DH;                      // heads per staged 128-byte row
  constexpr int QC = DH / 4;                        // accumulator columns per thread in the drain
  constexpr int NCHK = DH / 8;                      // 16-byte chunks of one head inside a staged row
  // TMEM columns: S^T | dP^T | dV | dK | dQ (tile 0); Dh = 32 keeps dQ of query tile 1 in the last 32 columns of S^T
  constexpr int COL_S = 0, COL_DP = DH == 32 ? 208 : 128, COL_DV = DH == 32 ? 416 : 256, COL_DK = COL_DV + DH, COL_DQ0 = COL_DK + DH;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic on the __shared__ array keeps the address space: LDS / STS, not generic LD / ST
  float* sL = reinterpret_cast<float*>(smem + Smem::OFF_L);
  float* sD = reinterpret_cast<float*>(smem + Smem::OFF_D);
  uint32_t* sRK = reinterpret_cast<uint32_t*>(smem + Smem::OFF_RK);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem::OFF_BAR);      // [0] S^T,dP^T  [1] dV,dK,dQ  [2] loads  [4..7] pass done, per warp group
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quad = warp & 3, grp = warp >> 2;
  const int row_in_tile = quad * 32 + lane;

  for (int i = tid; i < Smem::OFF_L / 16; i += THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    tc::prefetch_tmap(&tmQKV);
    tc::prefetch_tmap(&tmDO);
    tc::prefetch_tmap(&tmO);
    tc::mbar_init(&bars[0], 1); tc::mbar_init(&bars[1], 2); tc::mbar_init(&bars[2], 1);
    for (int g = 0; g < 4; ++g) tc::mbar_init(&bars[4 + g], 128);
    tc::fence_barrier_init();
  }
  if (warp == 0) tc::tmem_alloc(tmem_slot, 512);
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);

  const int ldq = 3 * H * DH;
  const int HG = H / HPB;
  const float c2 = scale * 1.4426950408889634f;
  uint32_t phase = 0, ph_ld = 0;
  const int n_items = ((nseq + pack - 1) / pack) * HG;
  const uint32_t sQ = tc::smem_u32(smem + Smem::OFF_Q), sDO = tc::smem_u32(smem + Smem::OFF_DO);
  const uint32_t sK = tc::smem_u32(smem + Smem::OFF_K), sV = tc::smem_u32(smem + Smem::OFF_V);
  const uint32_t sDS = tc::smem_u32(smem + Smem::OFF_DS);
  uint8_t* dsbuf = smem + Smem::OFF_DS;

  auto issue_loads = [&](int it) {                 // one thread
    const int sq = it / HG, hg = it - sq * HG;
    const int tt = seq_start[sq * pack], SS = seq_start[min(nseq, sq * pack + pack)] - tt;
    const int nb = (SS + BOX_ROWS - 1) / BOX_ROWS;
    tc::mbar_expect_tx(&bars[2], 5 * nb * BOX_BYTES);
    const int offs[3] = {Smem::OFF_Q, Smem::OFF_K, Smem::OFF_V};
#pragma unroll
    for (int m = 0; m < 3; ++m)
      for (int b = 0; b < nb; ++b)
        tc::tma_load_2d(smem + offs[m] + b * BOX_BYTES, &tmQKV, &bars[2], m * H * DH + hg * 64, tt + b * BOX_ROWS);
    for (int b = 0; b < nb; ++b) {
      tc::tma_load_2d(smem + Smem::OFF_DO + b * BOX_BYTES, &tmDO, &bars[2], hg * 64, tt + b * BOX_ROWS);
      tc::tma_load_2d(smem + Smem::OFF_O + b * BOX_BYTES, &tmO, &bars[2], hg * 64, tt + b * BOX_ROWS);
    }
  };
  if (tid == 0 && (int)blockIdx.x < n_items) issue_loads(blockIdx.x);
  if ((int)blockIdx.x < n_items) {                 // row statistics of the first item -> buffer 0
    const int sq = blockIdx.x / HG, hg0 = blockIdx.x - sq * HG;
    const int tt = seq_start[sq * pack], SS = seq_start[min(nseq, sq * pack + pack)] - tt;
    const int hd = tid >> 8, q = tid & 255;
    const bool ok = q < SS && hd < HPB;
    sL[tid] = ok ? lse[(size_t)(tt + q) * H + hg0 * HPB + hd] * 1.4426950408889634f : INFINITY;
  }

#ifdef EAVIT_TRACE
  long long tr_acc[16] = {0}, tr_prev = clock64();
#endif
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int seq = item / HG, hg = item - seq * HG;
    const int t0 = seq_start[seq * pack], S = seq_start[min(nseq, seq * pack + pack)] - t0;
    const int NQP = (S + 15) & ~15, nks = NQP >> 4;          // score columns (queries), k-steps of the dV / dK MMAs
    const int NT = (S + 127) >> 7;                           // key tiles == query tiles
    int kb[5];
    ksplit(nks, NT == 2, kb);
    const int cbeg = kb[grp] * 16, cend = kb[grp + 1] * 16;  // this thread's query columns
    const uint32_t col_dq1 = (uint32_t)(NQP - 32);           // dQ accumulator of query tile 1 (inside the last group's columns)
    // Row statistics of both heads of the pair, one entry per thread.  lse * log2(e) (+inf past the end: P = 0) is
    // double-buffered: the NEXT item's entry is fetched into a register now and parked in the other buffer after the
    // first key tile, so its global-load latency never sits in front of a round.  D = rowsum(O * dO) comes from the two
    // staged tiles (16-byte chunk c of row q sits at chunk c ^ (q & 7) of the swizzled row).
    const int sbuf = (int)(ph_ld & 1);
    tc::mbar_wait(&bars[2], ph_ld);
    ph_ld ^= 1;
    {
      const int hd = tid >> 8, q = tid & 255;
      float dsum = 0.f;
      if (q < S && hd < HPB) {
        const uint8_t* orow = smem + Smem::OFF_O + (q >> 3) * 1024 + (q & 7) * 128;
        const uint8_t* grow = smem + Smem::OFF_DO + (q >> 3) * 1024 + (q & 7) * 128;
#pragma unroll
        for (int j = 0; j < NCHK; ++j) {
          const int ch = ((hd * NCHK + j) ^ (q & 7)) << 4;
          const uint4 a = *reinterpret_cast<const uint4*>(orow + ch), b = *reinterpret_cast<const uint4*>(grow + ch);
          const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            dsum = fmaf(__uint_as_float(aw[i] << 16), __uint_as_float(bw[i] << 16), dsum);
            dsum = fmaf(__uint_as_float(aw[i] & 0xffff0000u), __uint_as_float(bw[i] & 0xffff0000u), dsum);
          }
        }
      }
      sD[tid] = dsum;
      if constexpr (DROP) {
        if (tid < 256) sRK[tid] = drop_row_key(drop, (uint32_t)(t0 + tid));
      }
    }
    __syncthreads();                                         // sL (written during the previous item) and sD visible
    const int nitem = item + (int)gridDim.x;
    float nl = INFINITY;
    if (nitem < n_items) {
      const int nsq = nitem / HG, nhg = nitem - nsq * HG;
      const int nt0 = seq_start[nsq * pack], nS = seq_start[min(nseq, nsq * pack + pack)] - nt0;
      const int hd = tid >> 8, q = tid & 255;
      if (q < nS && hd < HPB) nl = __ldg(lse + (size_t)(nt0 + q) * H + nhg * HPB + hd);      // raw: nothing depends on it until it is parked
    }
    BTR(0);

#pragma unroll 1
    for (int hd = 0; hd < HPB; ++hd) {
      const int h = hg * HPB + hd;
      const uint32_t hoff = (uint32_t)(hd * DH * 2);         // byte offset of this head inside the 128-byte rows
      const float* hL = sL + sbuf * 512 + hd * 256;
      const float* hD = sD + hd * 256;
      float accQ[2][QC];                                     // dQ rows (tile, row_in_tile), columns [grp*QC, +QC), summed over key tiles
#pragma unroll
      for (int i = 0; i < QC; ++i) { accQ[0][i] = 0.f; accQ[1][i] = 0.f; }

#pragma unroll 1
      for (int kt = 0; kt < NT; ++kt) {
        // ---- (1) S^T = K_kt Q^T, dP^T = V_kt dO^T     (M = 128 keys, N = NQP queries, K = Dh)
        if (warp == 0 && tc::elect_one()) {
          tc::fence_after_sync();
          const uint32_t idesc = tc::make_idesc_bf16(128, NQP, 0, 0);
          const uint64_t ak = tc::make_sdesc_sw128(sK + kt * 128 * ROWB + hoff, 16, 1024), bq = tc::make_sdesc_sw128(sQ + hoff, 16, 1024);
          const uint64_t av = tc::make_sdesc_sw128(sV + kt * 128 * ROWB + hoff, 16, 1024), bo = tc::make_sdesc_sw128(sDO + hoff, 16, 1024);
#pragma unroll
          for (int k = 0; k < DH / 16; ++k) tc::mma_bf16_ss(tmem_base + COL_S, ak + (uint64_t)(k * 2), bq + (uint64_t)(k * 2), idesc, k > 0);
#pragma unroll
          for (int k = 0; k < DH / 16; ++k) tc::mma_bf16_ss(tmem_base + COL_DP, av + (uint64_t)(k * 2), bo + (uint64_t)(k * 2), idesc, k > 0);
          tc::mma_commit(&bars[0]);
        }
        BTR(1);
        const int krow = kt * 128 + row_in_tile;             // this thread's key
        const bool kok = krow < S;
        const int nkeys = min(128, S - kt * 128);            // valid keys of the tile
        const int kq = (nkeys + 15) >> 4;                    // 16-key K steps of the dQ MMA
        const bool rows_live = kt * 128 + quad * 32 < S;     // warp-uniform: any valid key in this warp
        const bool all_ok = kt * 128 + quad * 32 + 31 < S;   // warp-uniform: every key of this warp is valid
        // dropout: one 32-bit hash serves the key pair (2j, 2j+1) of a query row; this thread's key picks its 16 bits
        // (mask column = key index inside the key's own sequence; seg is even, so pairs never straddle two sequences)
        int qlo = 0, qhi = S;                                // queries this key belongs to
        if constexpr (SEG) {
          if (kok) { qlo = (krow / seg) * seg; qhi = min(qlo + seg, S); }
        }
        const int kin = krow - qlo;
        const uint32_t ck = (uint32_t)(h * 128 + (kin >> 1)) * 0x9E3779B9u;
        const uint32_t ksh = (kin & 1) ? 16u : 0u;
        tc::mbar_wait(&bars[0], phase);
        BTR(2);
        tc::fence_after_sync();
        // ---- the pass: P~^T -> TMEM (over S^T), dS^T -> TMEM (over dP^T) and shared memory
        if (rows_live) {
          tc::tmem_stream16x2<8>(lane_base + COL_S, COL_DP - COL_S, cbeg, cend, [&](const uint32_t* rs, const uint32_t* rp, int c0) {
            const float4 l0 = *reinterpret_cast<const float4*>(hL + c0), l1 = *reinterpret_cast<const float4*>(hL + c0 + 4);
            const float4 d0 = *reinterpret_cast<const float4*>(hD + c0), d1 = *reinterpret_cast<const float4*>(hD + c0 + 4);
            const float lq[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
            const float dq[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
            uint32_t pu[4], du[4];
            if constexpr (DROP) {
              const uint4 k0 = *reinterpret_cast<const uint4*>(sRK + c0), k1 = *reinterpret_cast<const uint4*>(sRK + c0 + 4);
              const uint32_t rkq[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
#pragma unroll
              for (int j = 0; j < 8; j += 2) {
                float p0 = ex2_approx(fmaf(__uint_as_float(rs[j]), c2, -lq[j]));
                float p1 = ex2_approx(fmaf(__uint_as_float(rs[j + 1]), c2, -lq[j + 1]));
                if constexpr (SEG) {
                  p0 = (c0 + j >= qlo && c0 + j < qhi) ? p0 : 0.f;
                  p1 = (c0 + j + 1 >= qlo && c0 + j + 1 < qhi) ? p1 : 0.f;
                }
                const float m0 = ((lowbias32(rkq[j] + ck) >> ksh) & 0xffffu) >= drop.thresh ? drop.scale : 0.f;
                const float m1 = ((lowbias32(rkq[j + 1] + ck) >> ksh) & 0xffffu) >= drop.thresh ? drop.scale : 0.f;
                pu[j >> 1] = pack_bf16x2(p0 * m0, p1 * m1);
                du[j >> 1] = pack_bf16x2(p0 * fmaf(m0, __uint_as_float(rp[j]), -dq[j]), p1 * fmaf(m1, __uint_as_float(rp[j + 1]), -dq[j + 1]));
              }
            } else {
#pragma unroll
              for (int j = 0; j < 8; j += 2) {
                float p0 = ex2_approx(fmaf(__uint_as_float(rs[j]), c2, -lq[j]));
                float p1 = ex2_approx(fmaf(__uint_as_float(rs[j + 1]), c2, -lq[j + 1]));
                if constexpr (SEG) {
                  p0 = (c0 + j >= qlo && c0 + j < qhi) ? p0 : 0.f;
                  p1 = (c0 + j + 1 >= qlo && c0 + j + 1 < qhi) ? p1 : 0.f;
                }
                pu[j >> 1] = pack_bf16x2(p0, p1);
                du[j >> 1] = pack_bf16x2(p0 * (__uint_as_float(rp[j]) - dq[j]), p1 * (__uint_as_float(rp[j + 1]) - dq[j + 1]));
              }
            }
            if (!all_ok && !kok) {                           // a key past the end of the sequence: exact zeros (dQ sums over keys)
#pragma unroll
              for (int j = 0; j < 4; ++j) { pu[j] = 0u; du[j] = 0u; }
            }
            *reinterpret_cast<uint4*>(dsbuf + (c0 >> 6) * SLAB + sw_off(row_in_tile, (c0 & 63) >> 3)) = make_uint4(du[0], du[1], du[2], du[3]);
            // packed pairs go back over the first half of this thread's own columns: chunk i (columns cbeg + 8 i ..) lands
            // at cbeg + 4 i .., always behind the columns this thread still has to read
            const uint32_t half = (uint32_t)(cbeg + ((c0 - cbeg) >> 1));
            tc::tmem_st_32x4(lane_base + COL_S + half, pu);
            tc::tmem_st_32x4(lane_base + COL_DP + half, du);
          });
          tc::tmem_st_wait();
        }
        BTR(3);
        // this warp group's P~^T / dS^T columns are in TMEM: the issuer starts on them while slower groups still compute
        tc::fence_before_sync();
        tc::mbar_arrive(&bars[4 + grp]);
        // ---- (2) dV_kt = P~^T dO, dK_kt = dS^T Q  (A from TMEM, K = queries), group by group as they finish
        if (warp == 0 && tc::elect_one()) {
          const uint32_t idesc_t = tc::make_idesc_bf16(128, DH, 0, 1);      // A K-major (TMEM), B MN-major
          bool first = true;
          for (int g = 0; g < 4; ++g) {
            tc::mbar_wait(&bars[4 + g], phase);
            tc::fence_after_sync();
            for (int ks = kb[g]; ks < kb[g + 1]; ++ks) {
              const uint32_t a_col = (uint32_t)(kb[g] * 16 + (ks - kb[g]) * 8);
              const uint64_t bdo = tc::make_sdesc_sw128(sDO + ks * 2048 + hoff, 8192, 1024);
              const uint64_t bq = tc::make_sdesc_sw128(sQ + ks * 2048 + hoff, 8192, 1024);
              tc::mma_bf16_ts(tmem_base + COL_DV, tmem_base + COL_S + a_col, bdo, idesc_t, first ? 0u : 1u);
              tc::mma_bf16_ts(tmem_base + COL_DK, tmem_base + COL_DP + a_col, bq, idesc_t, first ? 0u : 1u);
              first = false;
            }
          }
          tc::mma_commit(&bars[1]);
        }
        tc::fence_proxy_async();
        __syncthreads();
        BTR(4);
        // ---- dQ_t += dS_t K_kt  (A = the complete dS^T buffer in shared memory, MN-major)
        if (warp == 4 && tc::elect_one()) {                                   // second issuer: the two run concurrently
          tc::fence_after_sync();
          const uint32_t idesc_q = tc::make_idesc_bf16(128, DH, 1, 1);      // A MN-major (queries), B MN-major
          for (int t = 0; t < NT; ++t) {
            const uint32_t dcol = t == 0 ? (uint32_t)COL_DQ0 : col_dq1;
            for (int kk = 0; kk < kq; ++kk) {
              const uint64_t a = tc::make_sdesc_sw128(sDS + 2 * t * SLAB + kk * 2048, SLAB, 1024);
              const uint64_t b = tc::make_sdesc_sw128(sK + (kt * 128 + kk * 16) * ROWB + hoff, 8192, 1024);
              tc::mma_bf16_ss(tmem_base + dcol, a, b, idesc_q, kk > 0);
            }
          }
          tc::mma_commit(&bars[1]);
        }
        BTR(5);
        tc::mbar_wait(&bars[1], phase);
        BTR(6);
        if (tid == 0 && hd == HPB - 1 && kt == NT - 1 && item + (int)gridDim.x < n_items)
          issue_loads(item + gridDim.x);            // every MMA on this item's operands is done: refill during the drain
        tc::fence_after_sync();
        // ---- drain: dV / dK rows of this key tile straight to global, dQ partials into registers
        {
          uint32_t rv[QC], rk[QC], q0[QC], q1[QC];
          tc::tmem_ld_w<QC>(lane_base + COL_DV + grp * QC, rv);
          tc::tmem_ld_w<QC>(lane_base + COL_DK + grp * QC, rk);
          tc::tmem_ld_w<QC>(lane_base + COL_DQ0 + grp * QC, q0);
          if (NT == 2) tc::tmem_ld_w<QC>(lane_base + col_dq1 + grp * QC, q1);
          tc::tmem_ld_wait();
          if (kok) {
            __nv_bfloat16* dk = dqkv + (size_t)(t0 + krow) * ldq + H * DH + h * DH + grp * QC;
            __nv_bfloat16* dv = dk + H * DH;
#pragma unroll
            for (int j = 0; j < QC; j += 8) {
              uint4 o;
              o.x = pack_bf16x2(__uint_as_float(rk[j]) * scale, __uint_as_float(rk[j + 1]) * scale);
              o.y = pack_bf16x2(__uint_as_float(rk[j + 2]) * scale, __uint_as_float(rk[j + 3]) * scale);
              o.z = pack_bf16x2(__uint_as_float(rk[j + 4]) * scale, __uint_as_float(rk[j + 5]) * scale);
              o.w = pack_bf16x2(__uint_as_float(rk[j + 6]) * scale, __uint_as_float(rk[j + 7]) * scale);
              *reinterpret_cast<uint4*>(dk + j) = o;
              o.x = pack_bf16x2(__uint_as_float(rv[j]), __uint_as_float(rv[j + 1])); o.y = pack_bf16x2(__uint_as_float(rv[j + 2]), __uint_as_float(rv[j + 3]));
              o.z = pack_bf16x2(__uint_as_float(rv[j + 4]), __uint_as_float(rv[j + 5])); o.w = pack_bf16x2(__uint_as_float(rv[j + 6]), __uint_as_float(rv[j + 7]));
              *reinterpret_cast<uint4*>(dv + j) = o;
            }
          }
#pragma unroll
          for (int i = 0; i < QC; ++i) {
            accQ[0][i] += __uint_as_float(q0[i]);
            if (NT == 2) accQ[1][i] += __uint_as_float(q1[i]);
          }
        }
        if (hd == 0 && kt == 0) sL[(sbuf ^ 1) * 512 + tid] = nl * 1.4426950408889634f;
        BTR(7);
        tc::fence_before_sync();
        phase ^= 1;
        __syncthreads();                              // TMEM and the dS buffer are free for the next key tile / head / item
        BTR(8);
      }
      // ---- dQ rows of this head
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int qrow = t * 128 + row_in_tile;
        if (t < NT && qrow < S) {
#pragma unroll
          for (int j = 0; j < QC; j += 8) {
            uint4 o;
            o.x = pack_bf16x2(accQ[t][j] * scale, accQ[t][j + 1] * scale); o.y = pack_bf16x2(accQ[t][j + 2] * scale, accQ[t][j + 3] * scale);
            o.z = pack_bf16x2(accQ[t][j + 4] * scale, accQ[t][j + 5] * scale); o.w = pack_bf16x2(accQ[t][j + 6] * scale, accQ[t][j + 7] * scale);
            *reinterpret_cast<uint4*>(dqkv + (size_t)(t0 + qrow) * ldq + h * DH + grp * QC + j) = o;
          }
        }
      }
    }
  }
#ifdef EAVIT_TRACE
  if (blockIdx.x == 0 && tid == 0)
    for (int i = 0; i < 16; ++i) g_bt_trace[i] += tr_acc[i];
#endif
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace bt
}  // namespace eavit

using namespace eavit;

template <int DH, bool DROP, bool SEG>
static int launch_bwd_t(const CUtensorMap& tq, const CUtensorMap& tdo, const CUtensorMap& to, const float* lse, const int* seq_start,
                        int nseq, int H, float scale, void* dqkv, const DropCfg& drop, int pack, int seg, cudaStream_t st) {
  static bool done = false;
  if (!done) {
    EAVIT_CUDA(cudaFuncSetAttribute(bt::attention_bwd_t_kernel<DH, DROP, SEG>, cudaFuncAttributeMaxDynamicSharedMemorySize, bt::Smem::TOTAL));
    done = true;
  }
  const int items = ((nseq + pack - 1) / pack) * (H / (64 / DH));
  const int grid = items < kNumSMs ? items : kNumSMs;
  bt::attention_bwd_t_kernel<DH, DROP, SEG><<<grid, bt::THREADS, bt::Smem::TOTAL, st>>>(tq, tdo, to, lse, seq_start, nseq, H, scale,
                                                                                        (__nv_bfloat16*)dqkv, drop, pack, seg);
  EAVIT_LAUNCH_OK();
  return EAVIT_OK;
}
template <int DH>
static int dispatch_bwd_t(const CUtensorMap& tq, const CUtensorMap& tdo, const CUtensorMap& to, const float* lse, const int* seq_start,
                          int nseq, int H, float scale, void* dqkv, const DropCfg& drop, int pack, int seg, cudaStream_t st) {
  if (pack > 1)
    return drop.thresh ? launch_bwd_t<DH, true, true>(tq, tdo, to, lse, seq_start, nseq, H, scale, dqkv, drop, pack, seg, st)
                       : launch_bwd_t<DH, false, true>(tq, tdo, to, lse, seq_start, nseq, H, scale, dqkv, drop, pack, seg, st);
  return drop.thresh ? launch_bwd_t<DH, true, false>(tq, tdo, to, lse, seq_start, nseq, H, scale, dqkv, drop, 1, 0, st)
                     : launch_bwd_t<DH, false, false>(tq, tdo, to, lse, seq_start, nseq, H, scale, dqkv, drop, 1, 0, st);
}

extern "C" int eavit_attention_bwd_tct(const void* qkv, const void* out, const void* dout, const float* lse,
                                       const int* seq_start, int nseq, int max_len, long long total_tokens, int H, int Dh,
                                       float scale, void* dqkv, float drop_p, unsigned long long drop_seed, void* stream) {
  EAVIT_CHECK_ARG(qkv && out && dout && lse && seq_start && dqkv && nseq > 0 && H > 0 && total_tokens > 0);
  EAVIT_CHECK_ARG((Dh == 32 && H % 2 == 0 && max_len <= bt::MAXQ) || (Dh == 64 && max_len <= bt::MAXQ64));
  EAVIT_CHECK_ARG(max_len > 0 && drop_p >= 0.f && drop_p < 1.f);
  EAVIT_CHECK_ARG((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(dout) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(out) & 15) == 0 && (reinterpret_cast<uintptr_t>(dqkv) & 15) == 0);
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap tq, tdo, to;
  int rc = make_tmap_bf16_2d(&tq, qkv, (uint64_t)3 * H * Dh, (uint64_t)total_tokens, (uint64_t)3 * H * Dh * 2, bt::BOX_ROWS);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tdo, dout, (uint64_t)H * Dh, (uint64_t)total_tokens, (uint64_t)H * Dh * 2, bt::BOX_ROWS);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&to, out, (uint64_t)H * Dh, (uint64_t)total_tokens, (uint64_t)H * Dh * 2, bt::BOX_ROWS);
  if (rc) return rc;
  const DropCfg drop = make_drop(drop_p, drop_seed);
  // sequences per work item: > 1 only when every sequence has the same even length and several fit into one item
  const int limit = Dh == 32 ? bt::MAXQ : bt::MAXQ64;
  int pack = 1;
  if (total_tokens == (long long)nseq * max_len && (max_len & 1) == 0 && 2 * max_len <= limit) pack = limit / max_len;
  if (Dh == 32) return dispatch_bwd_t<32>(tq, tdo, to, lse, seq_start, nseq, H, scale, dqkv, drop, pack, max_len, st);
  return dispatch_bwd_t<64>(tq, tdo, to, lse, seq_start, nseq, H, scale, dqkv, drop, pack, max_len, st);
}

#ifdef EAVIT_TRACE
extern "C" int eavit_debug_bt_trace(long long* host32, int reset) {
  if (reset) { long long z[32] = {0}; cudaMemcpyToSymbol(eavit::bt::g_bt_trace, z, sizeof(z)); return 0; }
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(host32, eavit::bt::g_bt_trace, 32 * sizeof(long long));
  return 0;
}
#endif
