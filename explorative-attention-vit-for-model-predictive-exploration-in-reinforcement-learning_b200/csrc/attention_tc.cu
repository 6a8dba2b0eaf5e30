// Tensor-core attention for short sequences (S <= 224) on tcgen05 / TMEM, one CTA (512 threads) per (sequence, head).
//
// The whole head fits on chip, so there is no online softmax and no K/V loop:
//   scores  S = Q K^T        tcgen05.mma  128 x NKT x DH   (A = Q tile, B = K, both K-major in 128B-swizzled smem)
//   softmax                  threads own (row, column-slice) pieces of the TMEM accumulator (tcgen05.ld), exp2 in fp32,
//                            un-normalised P written as bf16 into swizzled smem (the A operand of the next MMA)
//   output  O = P V          tcgen05.mma  128 x 64 x NKP   (B = V used in place as an MN-major operand -- no transpose)
//   epilogue                 O / rowsum -> bf16, lse = max + log(sum)
// The kernels are instruction-issue bound (exp2 / convert / pack per score; ncu: tensor pipe 7-11 %, DRAM 8 %), so each
// TMEM row is shared by several warps (the lane quadrant of warp w is w % 4 for every warp): 16 warps per SM instead of 8.
// Scores never leave the SM: the reference materialises [B,8,S,S] fp32 (vit.py:66-71).
#include "common.cuh"
#include "tc05.cuh"

namespace eavit {

constexpr int ATC_THREADS = 512;
constexpr int ATC_MAXKEYS = 224;          // TMEM: 2 tiles x 224 score columns = 448 <= 512
constexpr int ROWB = 128;                 // bytes per smem row (64 bf16): one 128B swizzle row
constexpr int P_SLAB = 128 * ROWB;        // [128 rows][64 keys]

// byte offset of 16-byte chunk `c` (0..7) of row `r` inside a K-major 128B-swizzled slab [rows][128 B]
__device__ __forceinline__ uint32_t sw_off(int r, int c) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

int make_tmap_bf16_2d(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_bytes, uint32_t box_outer);

// Operand staging: one elected thread issues TMA box loads (64 bf16 = 128 bytes x 32 rows, 128-byte swizzle) of the item's
// Q / K / V (/ dO) rows straight out of the [T, 3*H*Dh] activation matrix.  A 128-byte box row holds 64 / Dh heads, so a
// work item is (sequence, head group): with Dh = 32 two heads share one staged tile and are processed back to back (the
// second head is the +64-byte K-slice / N-slice of the same swizzled rows).  Rows past the end of the sequence inside the
// last 32-row box belong to the next sequence (or are zero-filled past T): finite values whose score columns are masked.
// Optional phase timing (-DEAVIT_TRACE, tools/att_trace.py): clock64 deltas of block 0 / thread 0 accumulated per phase.
#ifdef EAVIT_TRACE
__device__ long long g_att_trace[32];
#define TR(i) do { if (threadIdx.x == 0) { const long long _t = clock64(); tr_acc[i] += _t - tr_prev; tr_prev = _t; } } while (0)   // local accumulators: no global round trip inside the timed code
#else
#define TR(i) do { } while (0)
#endif
constexpr int BOX_ROWS = 32;
constexpr int BOX_BYTES = BOX_ROWS * ROWB;

struct AttSmem {
  static constexpr int OFF_Q = 0;                                // [256][128 B]
  static constexpr int OFF_K = OFF_Q + 256 * ROWB;               // [224][128 B]
  static constexpr int OFF_V = OFF_K + ATC_MAXKEYS * ROWB;
  static constexpr int OFF_P = OFF_V + ATC_MAXKEYS * ROWB;       // 2 tiles x 4 slabs
  static constexpr int OFF_X = OFF_P + 2 * 4 * P_SLAB;           // float [2 halves][256 rows] max, then sum
  static constexpr int OFF_BAR = OFF_X + 2 * 2 * 256 * 4;
  static constexpr int TOTAL = OFF_BAR + 64 + 1024;
};

// DROP: nn.Dropout on the attention probabilities (vit.py:45,70 / HF attention_probs_dropout_prob): the row sum (softmax
// denominator) uses the full probabilities, the P.V operand is masked.  Mask element = (token row t0 + q, column h*256 + k).
// SEG: short equal-length sequences are PACKED -- a work item is `pack` consecutive sequences (they are contiguous in the
// token matrix) treated as one sequence of pack * seg tokens under a block-diagonal mask: row i attends to the keys of
// its own segment [lo, hi) only.  Masked probabilities are exact zeros, so every later product (P.V, and dV / dK / dQ in
// the backward) is unchanged; the per-item latency chain (TMA -> MMA -> softmax -> MMA), which is all a 50-token
// sequence costs, is paid once per pack instead of once per sequence (HF-style ViT: 50 tokens -> 4 per item).
template <int DH, bool DROP, bool SEG>
__global__ void __launch_bounds__(ATC_THREADS, 1)
attention_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const int* __restrict__ seq_start, int nseq, int H,
                        float scale, __nv_bfloat16* __restrict__ out, float* __restrict__ lse, const DropCfg drop,
                        int pack, int seg) {
  constexpr int HPB = 64 / DH;                      // heads per 128-byte box row
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic on the __shared__ array keeps the address space: LDS / STS, not generic LD / ST
  float* sMax = reinterpret_cast<float*>(smem + AttSmem::OFF_X);      // [2][256]
  float* sSum = sMax + 512;                                            // [2][256]
  uint64_t* bar_s = reinterpret_cast<uint64_t*>(smem + AttSmem::OFF_BAR);
  uint64_t* bar_o = bar_s + 2;
  uint64_t* bar_qk = bar_s + 4;                  // Q, K staged
  uint64_t* bar_v = bar_s + 5;                   // V staged
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_s + 6);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quad = warp & 3, grp = warp >> 2;   // 4 groups of 4 warps
  const int g = grp & 1;                        // warp group (8 warps): owns TMEM region g, P buffer g, barriers g
  const int hf = grp >> 1;                      // column half of the tile's score row
  const int row_in_tile = quad * 32 + lane;
  const int xslot = g * 128 + row_in_tile;      // max / sum exchange slot (private to the group)

  for (int i = tid; i < AttSmem::OFF_X / 16; i += ATC_THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    tc::prefetch_tmap(&tmQKV);
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&bar_s[i], 1); tc::mbar_init(&bar_o[i], 1); }
    tc::mbar_init(bar_qk, 1);
    tc::mbar_init(bar_v, 1);
    tc::fence_barrier_init();
  }
  if (warp == 0) tc::tmem_alloc(tmem_slot, 512);
  tc::fence_proxy_async();                       // the zero fill above (generic proxy) precedes TMA writes to the same bytes
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  const int ldo = H * DH;
  const int HG = H / HPB;
  const float c2 = scale * 1.4426950408889634f;      // exp(x*scale) = exp2(x*c2)
  uint32_t ph0 = 0, ph1 = 0, ph_ld = 0;              // parities of bar_*[0], bar_*[1], bar_ld (uniform across the CTA)
  const int n_items = ((nseq + pack - 1) / pack) * HG;
  // Q and K are dead once the last head's score MMAs have completed, V once its P.V MMAs have: the next item's Q / K
  // land during the current item's last softmax and its V during the following one (TMA latency fully hidden).
  auto issue_qk = [&](int it) {                      // one thread
    const int sq = it / HG, hg = it - sq * HG;
    const int tt = seq_start[sq * pack], SS = seq_start[min(nseq, sq * pack + pack)] - tt;
    const int nb = (SS + BOX_ROWS - 1) / BOX_ROWS;
    tc::mbar_expect_tx(bar_qk, 2 * nb * BOX_BYTES);
    for (int b = 0; b < nb; ++b) {
      tc::tma_load_2d(smem + AttSmem::OFF_Q + b * BOX_BYTES, &tmQKV, bar_qk, hg * 64, tt + b * BOX_ROWS);
      tc::tma_load_2d(smem + AttSmem::OFF_K + b * BOX_BYTES, &tmQKV, bar_qk, H * DH + hg * 64, tt + b * BOX_ROWS);
    }
  };
  auto issue_v = [&](int it) {
    const int sq = it / HG, hg = it - sq * HG;
    const int tt = seq_start[sq * pack], SS = seq_start[min(nseq, sq * pack + pack)] - tt;
    const int nb = (SS + BOX_ROWS - 1) / BOX_ROWS;
    tc::mbar_expect_tx(bar_v, nb * BOX_BYTES);
    for (int b = 0; b < nb; ++b)
      tc::tma_load_2d(smem + AttSmem::OFF_V + b * BOX_BYTES, &tmQKV, bar_v, 2 * H * DH + hg * 64, tt + b * BOX_ROWS);
  };
  if (tid == 0 && (int)blockIdx.x < n_items) { issue_qk(blockIdx.x); issue_v(blockIdx.x); }
#ifdef EAVIT_TRACE
  long long tr_acc[20] = {0}, tr_prev = clock64();
#endif
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int seq = item / HG, hg = item - seq * HG;
    const int t0 = seq_start[seq * pack], S = seq_start[min(nseq, seq * pack + pack)] - t0;
    const int NKT = (S + 31) & ~31;                  // score columns
    const int NKP = (S + 15) & ~15;                  // keys covered by the P.V MMA
    const bool two_tiles = S > 128;
    const int nt = two_tiles ? 2 : 1;
    tc::mbar_wait(bar_qk, ph_ld);
    TR(0);
    const bool has_next = item + (int)gridDim.x < n_items;
    // Work split.  Dh = 32 (two heads per staged row): group g owns head g and walks its query tiles, group 0 upwards
    // and group 1 downwards -- with 196/197 tokens one group is in its long (128-row) tile while the other is in its
    // short one, so their MMA round trips and softmax phases interleave instead of running in lock step.
    // Dh = 64 (one head per item): group g owns query tile g.
    const int nsub0 = HPB == 2 ? nt : 1, nsub1 = HPB == 2 ? nt : (nt - 1);
    const int nsub = g ? nsub1 : nsub0;
    const int hd = HPB == 2 ? g : 0;
    const int h = hg * HPB + hd;
    const uint32_t hoff = (uint32_t)(hd * DH * 2);     // byte offset of this head inside the 128-byte rows
    const int v_issuer_grp = nsub1 > 0 ? 1 : 0;        // the group that finishes last refills V (group 1 ends on its long tile)
#pragma unroll 1
    for (int j = 0; j < nsub; ++j) {
      const int tl = HPB == 2 ? (g ? nt - 1 - j : j) : g;      // query tile of this sub-item
      const int xrow = tl * 128 + row_in_tile;
      const bool last = j == nsub - 1;
      int lo = 0, hi = S;                              // keys this row attends to
      if constexpr (SEG) {
        if (xrow < S) { lo = (xrow / seg) * seg; hi = min(lo + seg, S); }     // rows past the end keep [0, S): finite garbage, never stored
      }
      {
        const uint32_t phase = (g ? ph1 : ph0) ^ (uint32_t)(j & 1);
        const uint32_t s_col = tmem_base + (uint32_t)(g * ATC_MAXKEYS);
        // Dh = 32: P~ never leaves tensor memory -- pass 2 writes it back as packed bf16 pairs over the first half of each
        // thread's own score columns (tcgen05.st) and O = P~ V takes A from TMEM; the O accumulator sits in the 64 spare
        // columns behind the two score regions.  Dh = 64 (64-column accumulator) keeps P~ in shared memory and O over S.
        constexpr bool PT = DH == 32;
        const uint32_t o_col = PT ? tmem_base + (uint32_t)(2 * ATC_MAXKEYS + g * DH) : s_col;
        const bool issuer = (hf == 0 && quad == 0 && lane == 0);
        if (issuer) {
          tc::fence_after_sync();
          const uint32_t idesc = tc::make_idesc_bf16(128, NKT, 0, 0);
          const uint32_t qa = tc::smem_u32(smem + AttSmem::OFF_Q + tl * 128 * ROWB) + hoff;
          const uint32_t ka = tc::smem_u32(smem + AttSmem::OFF_K) + hoff;
          const uint64_t adesc = tc::make_sdesc_sw128(qa, 16, 1024), bdesc = tc::make_sdesc_sw128(ka, 16, 1024);
#pragma unroll
          for (int k = 0; k < DH / 16; ++k) tc::mma_bf16_ss(s_col, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, k > 0);
          tc::mma_commit(&bar_s[g]);
        }
        TR(1);
        tc::mbar_wait(&bar_s[g], phase);
        TR(2);
        if (tid == 0 && last && has_next) {
          // Q / K are dead once EVERY score MMA of the item is complete: the other group's last one as well
          if (nsub1 > 0) tc::mbar_wait(&bar_s[1], ph1 ^ (uint32_t)((nsub1 - 1) & 1));
          issue_qk(item + gridDim.x);
        }
        tc::fence_after_sync();
        const uint32_t lane_addr = s_col + ((uint32_t)(quad * 32) << 16);
        const int half = NKT >> 1, cbeg = hf * half;      // half is a multiple of 16
        // Warps whose 32 query rows all lie beyond the sequence (tile 1 of a 196/197-token sequence: rows >= 224) only keep
        // the barriers company; their P rows are row-local garbage that never reaches a stored output row.
        const bool rows_live = tl * 128 + quad * 32 < S;
        // pass 1: partial row max over this thread's columns (TMEM loads software-pipelined against the compares)
        float mx = -INFINITY;
        if (rows_live) {
          tc::tmem_stream16(lane_addr, cbeg, cbeg + half, [&](const uint32_t* r, int c0) {
            if (SEG && (c0 + 16 <= lo || c0 >= hi)) return;      // another sequence's keys
            if (SEG ? (c0 >= lo && c0 + 16 <= hi) : (c0 + 16 <= S)) {
              // four independent max chains (a single one serialises 112 dependent FMNMX per row)
              float m0 = fmaxf(__uint_as_float(r[0]), __uint_as_float(r[1])), m1 = fmaxf(__uint_as_float(r[2]), __uint_as_float(r[3]));
              float m2 = fmaxf(__uint_as_float(r[4]), __uint_as_float(r[5])), m3 = fmaxf(__uint_as_float(r[6]), __uint_as_float(r[7]));
              m0 = fmaxf(m0, fmaxf(__uint_as_float(r[8]), __uint_as_float(r[9]))); m1 = fmaxf(m1, fmaxf(__uint_as_float(r[10]), __uint_as_float(r[11])));
              m2 = fmaxf(m2, fmaxf(__uint_as_float(r[12]), __uint_as_float(r[13]))); m3 = fmaxf(m3, fmaxf(__uint_as_float(r[14]), __uint_as_float(r[15])));
              mx = fmaxf(mx, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (c0 + j >= lo && c0 + j < hi) mx = fmaxf(mx, __uint_as_float(r[j]));
            }
          });
        }
        sMax[hf * 256 + xslot] = mx;
        TR(3);
        named_bar_sync(1 + g, 256);
        TR(4);
        mx = fmaxf(sMax[xslot], sMax[256 + xslot]);
        // pass 2: p = exp2((s - mx) * c2) -> bf16 P tile in swizzled smem; the row sum is taken in fp32 before rounding
        // (|sum(p~) - sum(p)| / sum(p) ~ 2^-9 / sqrt(S), far below the bf16 rounding of the output)
        const float mb = mx * c2;
        float sum = 0.f;
        uint8_t* pbase = smem + AttSmem::OFF_P + g * 4 * P_SLAB;
        uint32_t rk = 0;
        if constexpr (DROP) rk = drop_row_key(drop, (uint32_t)(t0 + xrow));
        const uint32_t hcol = (uint32_t)h * 256u;
        if (rows_live) {
          tc::tmem_stream16(lane_addr, cbeg, cbeg + half, [&](const uint32_t* r, int c0) {
            if (c0 >= NKP) return;                       // beyond the keys the P.V MMA reads
            uint8_t* slab = pbase + (c0 >> 6) * P_SLAB;
            const int cb = (c0 & 63) >> 3;
            uint32_t pk[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
            const uint32_t dcol = hcol + (uint32_t)(c0 - lo);      // dropout column = key index inside the row's own sequence (lo is even)
            // (SEG: the lanes of a warp sit in different sequences -- every path falls through to the common tail, whose
            // TMEM store is warp-collective)
            if (SEG && (c0 + 16 <= lo || c0 >= hi)) {
              // another sequence's keys: exact zeros, no math
            } else if (SEG ? (c0 >= lo && c0 + 16 <= hi) : (c0 + 16 <= S)) {
              float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
              for (int j = 0; j < 16; j += 2) {
                float p0 = ex2_approx(fmaf(__uint_as_float(r[j]), c2, -mb));
                float p1 = ex2_approx(fmaf(__uint_as_float(r[j + 1]), c2, -mb));
                if (j & 2) { s2 += p0; s3 += p1; } else { s0 += p0; s1 += p1; }
                if constexpr (DROP) {
                  const uint32_t bits = drop_bits(rk, (dcol + (uint32_t)j) >> 1);
                  p0 *= drop_even(drop, bits); p1 *= drop_odd(drop, bits);
                }
                pk[j >> 1] = pack_bf16x2(p0, p1);
              }
              sum += (s0 + s1) + (s2 + s3);
            } else {
#pragma unroll
              for (int j = 0; j < 16; j += 2) {
                float p0 = (c0 + j >= lo && c0 + j < hi) ? ex2_approx(fmaf(__uint_as_float(r[j]), c2, -mb)) : 0.f;
                float p1 = (c0 + j + 1 >= lo && c0 + j + 1 < hi) ? ex2_approx(fmaf(__uint_as_float(r[j + 1]), c2, -mb)) : 0.f;
                sum += p0 + p1;
                if constexpr (DROP) {
                  const uint32_t bits = drop_bits(rk, (dcol + (uint32_t)j) >> 1);
                  p0 *= drop_even(drop, bits); p1 *= drop_odd(drop, bits);
                }
                pk[j >> 1] = pack_bf16x2(p0, p1);
              }
            }
            if constexpr (PT) {
              // chunk i (columns cbeg + 16 i ..) lands at cbeg + 8 i ..: always behind the columns this thread still reads
              if constexpr (SEG) __syncwarp();
              tc::tmem_st_32x8(lane_addr + (uint32_t)(cbeg + ((c0 - cbeg) >> 1)), pk);
            } else {
              *reinterpret_cast<uint4*>(slab + sw_off(row_in_tile, cb)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              *reinterpret_cast<uint4*>(slab + sw_off(row_in_tile, cb + 1)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            }
          });
          if constexpr (PT) tc::tmem_st_wait();
        }
        sSum[hf * 256 + xslot] = sum;
        TR(5);
        // all S reads done (Dh = 64: O overlays S) and P visible to the tensor core, then one thread issues P.V
        tc::fence_before_sync();
        if constexpr (!PT) tc::fence_proxy_async();
        named_bar_sync(1 + g, 256);
        TR(6);
        if (issuer) {
          if (j == 0) tc::mbar_wait(bar_v, ph_ld);
          tc::fence_after_sync();
          const uint32_t idesc = tc::make_idesc_bf16(128, DH, 0, 1);
          const uint32_t pa = tc::smem_u32(pbase);
          const uint32_t va = tc::smem_u32(smem + AttSmem::OFF_V) + hoff;    // N-slice of this head inside the MN-major rows
          for (int kk = 0; kk < NKP / 16; ++kk) {
            const uint64_t bdesc = tc::make_sdesc_sw128(va + kk * 2048, 8192, 1024);
            if constexpr (PT) {
              const int c16 = kk * 16;                       // packed pairs of columns [c16, +16): 8 TMEM columns inside their half
              const uint32_t a_col = (uint32_t)(c16 < half ? (c16 >> 1) : half + ((c16 - half) >> 1));
              tc::mma_bf16_ts(o_col, s_col + a_col, bdesc, idesc, kk > 0);
            } else {
              const uint64_t adesc = tc::make_sdesc_sw128(pa + (kk >> 2) * P_SLAB + (kk & 3) * 32, 16, 1024);
              tc::mma_bf16_ss(s_col, adesc, bdesc, idesc, kk > 0);
            }
          }
          tc::mma_commit(&bar_o[g]);
        }
        sum = sSum[xslot] + sSum[256 + xslot];
        TR(7);
        tc::mbar_wait(&bar_o[g], phase);
        TR(8);
        if (issuer && g == v_issuer_grp && last && has_next) {
          // V is dead once every P.V MMA of the item is complete: wait for the other group's last one, then refill
          const int og = g ^ 1, onsub = og ? nsub1 : nsub0;
          if (onsub > 0) tc::mbar_wait(&bar_o[og], (og ? ph1 : ph0) ^ (uint32_t)((onsub - 1) & 1));
          issue_v(item + gridDim.x);
        }
        tc::fence_after_sync();
        if (rows_live) {
          constexpr int HW = DH / 2;                   // output columns per thread
          uint32_t r[32];
          const uint32_t o_lane = o_col + ((uint32_t)(quad * 32) << 16);
          if constexpr (HW == 16) tc::tmem_ld_32x16(o_lane + hf * HW, r);
          else tc::tmem_ld_32x32(o_lane + hf * HW, r);
          tc::tmem_ld_wait();
          const int qrow = xrow;
          const float inv = 1.f / sum;
          if (qrow < S) {
            __nv_bfloat16* orow = out + (size_t)(t0 + qrow) * ldo + h * DH + hf * HW;
#pragma unroll
            for (int j = 0; j < HW; j += 8) {
              uint4 o;
              o.x = pack_bf16x2(__uint_as_float(r[j]) * inv, __uint_as_float(r[j + 1]) * inv);
              o.y = pack_bf16x2(__uint_as_float(r[j + 2]) * inv, __uint_as_float(r[j + 3]) * inv);
              o.z = pack_bf16x2(__uint_as_float(r[j + 4]) * inv, __uint_as_float(r[j + 5]) * inv);
              o.w = pack_bf16x2(__uint_as_float(r[j + 6]) * inv, __uint_as_float(r[j + 7]) * inv);
              *reinterpret_cast<uint4*>(orow + j) = o;
            }
            if (hf == 0 && lse != nullptr) lse[(size_t)(t0 + qrow) * H + h] = mx * scale + __logf(sum);
          }
        }
        tc::fence_before_sync();
      }
      // TMEM (O overlays S), the P buffer and the max / sum exchange slots are private to the group: only the group
      // has to agree that they are free for the next sub-item / item
      TR(9);
      named_bar_sync(1 + g, 256);
      TR(10);
    }
    ph0 ^= (uint32_t)(nsub0 & 1);
    ph1 ^= (uint32_t)(nsub1 & 1);
    ph_ld ^= 1;
  }
#ifdef EAVIT_TRACE
  if (blockIdx.x == 0 && tid == 0)
    for (int i = 0; i < 20; ++i) g_att_trace[i] += tr_acc[i];
#endif
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, 512);
  }
}

// =====================================================================================================================
// Backward.  Per (sequence, head) item and per 128-row query tile g, all 512 threads work on one tile: the four groups
// of four warps split the key columns of each TMEM row four ways.
//   S  = Q_g K^T, dP = dO_g V^T                tcgen05.mma 128 x NKT x DH  (TMEM cols [0,NKT) and [224,224+NKT))
//   P~ = bf16(exp2(S c2 - lse log2e))           -> swizzled smem;   D_i = sum_j P_ij dP_ij with the fp32 P (sum_j P_ij = 1 to
//                                                  fp32 accuracy, so dS = P~ (dP - D) carries no D-proportional bias)
//   dV_kt += P~^T dO_g                          tcgen05.mma 128 x 64 x 128, A = P~ read in place as an MN-major operand
//   dS = P~ (dP - D)                            -> the same smem buffer once dV has consumed P~
//   dK_kt += dS^T Q_g ; dQ_g = dS K             A = dS MN-major / K-major, B = Q_g / K in place as MN-major operands
// dK / dV partial sums of the two query tiles are accumulated in registers (half a key row per thread).
// =====================================================================================================================
struct AttBwdSmem {
  static constexpr int OFF_Q = 0;                                  // [256][128 B]
  static constexpr int OFF_DO = OFF_Q + 256 * ROWB;                // [256][128 B]
  static constexpr int OFF_K = OFF_DO + 256 * ROWB;                // [224][128 B]
  static constexpr int OFF_V = OFF_K + ATC_MAXKEYS * ROWB;
  static constexpr int OFF_P = OFF_V + ATC_MAXKEYS * ROWB;         // 4 slabs [128][64 keys]
  static constexpr int OFF_D = OFF_P + 4 * P_SLAB;                 // float [4][128]
  static constexpr int OFF_BAR = OFF_D + 4 * 128 * 4;
  static constexpr int TOTAL = OFF_BAR + 64 + 1024;
};

// DROP: with dropout mask M on the probabilities, dV = (P o M)^T dO, D_i = sum_j P_ij M_ij dP_ij and dS = P o (M o dP - D):
// pass 1 stages the MASKED P~ for the dV MMA; pass 2 needs the unmasked P again and recomputes it from S, which stays
// intact because the dV accumulators live in the last 64 TMEM columns instead of overlaying S.
template <int DH, bool DROP, bool SEG>
__global__ void __launch_bounds__(ATC_THREADS, 1)
attention_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                        const float* __restrict__ lse, const int* __restrict__ seq_start, int nseq, int H, float scale,
                        __nv_bfloat16* __restrict__ dqkv, const DropCfg drop, int pack, int seg) {
  constexpr int HPB = 64 / DH;                      // heads per 128-byte box row
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic on the __shared__ array keeps the address space: LDS / STS, not generic LD / ST
  float* sD = reinterpret_cast<float*>(smem + AttBwdSmem::OFF_D);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AttBwdSmem::OFF_BAR);   // [0] S,dP  [1] dV  [2] dK,dQ  [3] loads
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quad = warp & 3, grp = warp >> 2;      // 4 column groups
  const int row_in_tile = quad * 32 + lane;
  const int kt = grp & 1, kh = grp >> 1;           // key tile / column half owned for the dK, dV accumulators

  for (int i = tid; i < AttBwdSmem::OFF_D / 16; i += ATC_THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    tc::prefetch_tmap(&tmQKV);
    tc::prefetch_tmap(&tmDO);
    for (int i = 0; i < 4; ++i) tc::mbar_init(&bars[i], 1);
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
  constexpr int COL_DP = 224, COL_DV = 448, COL_DK = 0, COL_DQ = 128;   // dV: 2 x Dh (Dh = 32) or 1 x 64 columns at 448
  constexpr int HC = DH / 2;                       // accumulator columns per thread
  uint32_t phase = 0, ph_ld = 0;
  const int n_items = ((nseq + pack - 1) / pack) * HG;
  const uint32_t sQ = tc::smem_u32(smem + AttBwdSmem::OFF_Q), sDO = tc::smem_u32(smem + AttBwdSmem::OFF_DO);
  const uint32_t sK = tc::smem_u32(smem + AttBwdSmem::OFF_K), sV = tc::smem_u32(smem + AttBwdSmem::OFF_V);
  const uint32_t sP = tc::smem_u32(smem + AttBwdSmem::OFF_P);
  uint8_t* pbuf = smem + AttBwdSmem::OFF_P;

  auto issue_loads = [&](int it) {                 // one thread
    const int sq = it / HG, hg = it - sq * HG;
    const int tt = seq_start[sq * pack], SS = seq_start[min(nseq, sq * pack + pack)] - tt;
    const int nb = (SS + BOX_ROWS - 1) / BOX_ROWS;
    tc::mbar_expect_tx(&bars[3], 4 * nb * BOX_BYTES);
    const int offs[3] = {AttBwdSmem::OFF_Q, AttBwdSmem::OFF_K, AttBwdSmem::OFF_V};
#pragma unroll
    for (int m = 0; m < 3; ++m)
      for (int b = 0; b < nb; ++b)
        tc::tma_load_2d(smem + offs[m] + b * BOX_BYTES, &tmQKV, &bars[3], m * H * DH + hg * 64, tt + b * BOX_ROWS);
    for (int b = 0; b < nb; ++b)
      tc::tma_load_2d(smem + AttBwdSmem::OFF_DO + b * BOX_BYTES, &tmDO, &bars[3], hg * 64, tt + b * BOX_ROWS);
  };
  if (tid == 0 && (int)blockIdx.x < n_items) issue_loads(blockIdx.x);
#ifdef EAVIT_TRACE
  long long tr_acc[20] = {0}, tr_prev = 0;
#endif
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int seq = item / HG, hg = item - seq * HG;
    const int t0 = seq_start[seq * pack], S = seq_start[min(nseq, seq * pack + pack)] - t0;
    const int NKP = (S + 15) & ~15, NKT = NKP;     // score columns = keys covered by the dQ MMA (multiple of 16: 208 at S = 196 / 197)
    const int NQ = (S + 127) >> 7;                 // query tiles == key tiles
#ifdef EAVIT_TRACE
    tr_prev = clock64();
#endif
    tc::mbar_wait(&bars[3], ph_ld);
    TR(0);
    ph_ld ^= 1;

#pragma unroll 1
    for (int hd = 0; hd < HPB; ++hd) {
    const int h = hg * HPB + hd;
    const uint32_t hoff = (uint32_t)(hd * DH * 2);   // byte offset of this head inside the 128-byte rows
    float accK[HC], accV[HC];                      // key row kt*128 + row_in_tile, columns [kh*HC, +HC), summed over query tiles
#pragma unroll
    for (int d = 0; d < HC; ++d) { accK[d] = 0.f; accV[d] = 0.f; }

    for (int g = 0; g < NQ; ++g) {
      // ---- (1) S = Q_g K^T, (2) dP = dO_g V^T
      if (warp == 0 && tc::elect_one()) {      // warp-uniform election: the MMAs issue without a per-lane replay loop
        tc::fence_after_sync();
        TR(13);
        const uint32_t idesc = tc::make_idesc_bf16(128, NKT, 0, 0);
        const uint64_t aq = tc::make_sdesc_sw128(sQ + g * 128 * ROWB + hoff, 16, 1024), bk = tc::make_sdesc_sw128(sK + hoff, 16, 1024);
        const uint64_t ao = tc::make_sdesc_sw128(sDO + g * 128 * ROWB + hoff, 16, 1024), bv = tc::make_sdesc_sw128(sV + hoff, 16, 1024);
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) tc::mma_bf16_ss(tmem_base, aq + (uint64_t)(k * 2), bk + (uint64_t)(k * 2), idesc, k > 0);
        TR(14);
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) tc::mma_bf16_ss(tmem_base + COL_DP, ao + (uint64_t)(k * 2), bv + (uint64_t)(k * 2), idesc, k > 0);
        TR(15);
        tc::mma_commit(&bars[0]);
      }
      TR(1);
      const int qrow = g * 128 + row_in_tile;
      const bool qok = qrow < S;
      int lo = 0, hi = S;                            // keys of this row's own sequence (SEG: packed sequences, block-diagonal mask)
      if constexpr (SEG) {
        if (qok) { lo = (qrow / seg) * seg; hi = min(lo + seg, S); }
      }
      const int kq = (min(128, S - g * 128) + 15) >> 4;   // 16-row K steps of the dK / dV MMAs that hold valid queries
      const float l2 = qok ? lse[(size_t)(t0 + qrow) * H + h] * 1.4426950408889634f : INFINITY;   // invalid row: P~ = 0
      tc::mbar_wait(&bars[0], phase);
      TR(2);
      tc::fence_after_sync();
      // the four warp groups split the NKT/8 column chunks of each row evenly (7 chunks of 8 at NKT = 224)
      const int nch = NKT >> 3;
      const int cbeg = ((grp * nch) >> 2) << 3, cend = (((grp + 1) * nch) >> 2) << 3;
      const bool rows_live = g * 128 + quad * 32 < S;   // warp-uniform: any valid query row in this warp
      // ---- pass 1: P~ -> smem, partial D  (TMEM loads of chunk i+1 in flight while chunk i is processed)
      float dpart = 0.f;
      uint32_t rk = 0;
      if constexpr (DROP) rk = drop_row_key(drop, (uint32_t)(t0 + qrow));
      const uint32_t hcol = (uint32_t)h * 256u;
      if (rows_live) {
        tc::tmem_stream16x2<8>(lane_base, COL_DP, cbeg, cend, [&](const uint32_t* rs, const uint32_t* rp, int c0) {
          uint32_t pk[4] = {0u, 0u, 0u, 0u};
          uint32_t pu[4] = {0u, 0u, 0u, 0u};               // DROP: the unmasked P~ of this chunk (see below)
          float d0 = 0.f, d1 = 0.f;
          const uint32_t dcol = hcol + (uint32_t)(c0 - lo);
          // (lanes of a warp sit in different sequences when SEG: every path below falls through to the common tail, the
          // TMEM store there is warp-collective)
          if (SEG && (c0 + 8 <= lo || c0 >= hi)) {
            // another sequence's keys: P~ = 0 (and dS stays 0 in pass 2)
          } else if (SEG ? (c0 >= lo && c0 + 8 <= hi) : (c0 + 8 <= S)) {
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
              float p0 = ex2_approx(fmaf(__uint_as_float(rs[j]), c2, -l2));
              float p1 = ex2_approx(fmaf(__uint_as_float(rs[j + 1]), c2, -l2));
              if constexpr (DROP) {
                pu[j >> 1] = pack_bf16x2(p0, p1);
                const uint32_t bits = drop_bits(rk, (dcol + (uint32_t)j) >> 1);
                p0 *= drop_even(drop, bits); p1 *= drop_odd(drop, bits);
              }
              pk[j >> 1] = pack_bf16x2(p0, p1);
              d0 = fmaf(p0, __uint_as_float(rp[j]), d0);
              d1 = fmaf(p1, __uint_as_float(rp[j + 1]), d1);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
              float p0 = (c0 + j >= lo && c0 + j < hi) ? ex2_approx(fmaf(__uint_as_float(rs[j]), c2, -l2)) : 0.f;
              float p1 = (c0 + j + 1 >= lo && c0 + j + 1 < hi) ? ex2_approx(fmaf(__uint_as_float(rs[j + 1]), c2, -l2)) : 0.f;
              if constexpr (DROP) {
                pu[j >> 1] = pack_bf16x2(p0, p1);
                const uint32_t bits = drop_bits(rk, (dcol + (uint32_t)j) >> 1);
                p0 *= drop_even(drop, bits); p1 *= drop_odd(drop, bits);
              }
              pk[j >> 1] = pack_bf16x2(p0, p1);
              d0 = fmaf(p0, __uint_as_float(rp[j]), d0);
              d1 = fmaf(p1, __uint_as_float(rp[j + 1]), d1);
            }
          }
          dpart += d0 + d1;
          *reinterpret_cast<uint4*>(pbuf + (c0 >> 6) * P_SLAB + sw_off(row_in_tile, (c0 & 63) >> 3)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          // DROP: pass 2 needs the UNMASKED probabilities (dS = P (M dP - D) also for the dropped entries).  They go back
          // into TMEM as packed bf16 pairs, over the first half of this thread's own score columns: chunk i (score columns
          // cbeg + 8 i ..) lands at cbeg + 4 i .., always behind the columns this thread still has to read, and no other
          // thread touches this row's [cbeg, cend).  Pass 2 then costs one half-width TMEM read instead of re-reading S,
          // re-evaluating exp2 and re-hashing the mask.
          if constexpr (DROP) {
            __syncwarp();
            tc::tmem_st_32x4(lane_base + (uint32_t)(cbeg + ((c0 - cbeg) >> 1)), pu);
          }
        });
        if constexpr (DROP) tc::tmem_st_wait();
      } else {
        // no valid query row in this warp: its P~ / dS rows are zero (they feed the dK / dV sums over queries)
        for (int c0 = cbeg; c0 < cend; c0 += 8)
          *reinterpret_cast<uint4*>(pbuf + (c0 >> 6) * P_SLAB + sw_off(row_in_tile, (c0 & 63) >> 3)) = make_uint4(0, 0, 0, 0);
      }
      sD[grp * 128 + row_in_tile] = dpart;
      TR(3);
      tc::fence_before_sync();
      tc::fence_proxy_async();
      __syncthreads();
      TR(4);
      // ---- (3) dV_kt = P~^T dO_g   (accumulators in the spare TMEM columns 448..511)
      if (warp == 0 && tc::elect_one()) {      // warp-uniform election: the MMAs issue without a per-lane replay loop
        tc::fence_after_sync();
        const uint32_t idesc = tc::make_idesc_bf16(128, DH, 1, 1);
        for (int t = 0; t < NQ; ++t)
          for (int kk = 0; kk < kq; ++kk) {          // K dimension = the valid query rows of this tile only (the rest of P~ is zero)
            const uint64_t a = tc::make_sdesc_sw128(sP + 2 * t * P_SLAB + kk * 2048, P_SLAB, 1024);
            const uint64_t b = tc::make_sdesc_sw128(sDO + g * 128 * ROWB + kk * 2048 + hoff, 8192, 1024);
            tc::mma_bf16_ss(tmem_base + COL_DV + t * DH, a, b, idesc, kk > 0);
          }
        tc::mma_commit(&bars[1]);
      }
      TR(5);
      const float Di = (sD[row_in_tile] + sD[128 + row_in_tile]) + (sD[256 + row_in_tile] + sD[384 + row_in_tile]);
      tc::mbar_wait(&bars[1], phase);             // dV has consumed P~: the buffer can take dS
      TR(6);
      tc::fence_after_sync();
      // ---- pass 2: dS = P~ (dP - D) -> smem
      if (rows_live && DROP) {
        // unmasked P~ from TMEM (stashed by pass 1), the keep decision from the masked P~ still in shared memory
        // (zero = dropped, or a probability that rounded to zero -- dS is zero either way)
        tc::tmem_stream8_half(lane_base + COL_DP, lane_base + (uint32_t)cbeg, cbeg, cend, [&](const uint32_t* rp, const uint32_t* pp, int c0) {
          uint8_t* addr = pbuf + (c0 >> 6) * P_SLAB + sw_off(row_in_tile, (c0 & 63) >> 3);
          const uint4 pm = *reinterpret_cast<uint4*>(addr);
          const uint32_t pmw[4] = {pm.x, pm.y, pm.z, pm.w};
          uint32_t ds[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float px = __uint_as_float(pp[j] << 16), py = __uint_as_float(pp[j] & 0xffff0000u);
            const float mx = (pmw[j] & 0xffffu) ? drop.scale : 0.f, my = (pmw[j] >> 16) ? drop.scale : 0.f;
            ds[j] = pack_bf16x2(px * fmaf(mx, __uint_as_float(rp[2 * j]), -Di), py * fmaf(my, __uint_as_float(rp[2 * j + 1]), -Di));
          }
          *reinterpret_cast<uint4*>(addr) = make_uint4(ds[0], ds[1], ds[2], ds[3]);
        });
      } else if (rows_live) {
        tc::tmem_stream16<8>(lane_base + COL_DP, cbeg, cend, [&](const uint32_t* rp, int c0) {
          if (SEG && (c0 + 8 <= lo || c0 >= hi)) return;
          uint8_t* addr = pbuf + (c0 >> 6) * P_SLAB + sw_off(row_in_tile, (c0 & 63) >> 3);
          const uint4 pa = *reinterpret_cast<uint4*>(addr);
          const uint32_t pin[4] = {pa.x, pa.y, pa.z, pa.w};
          uint32_t ds[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float px = __uint_as_float(pin[j] << 16), py = __uint_as_float(pin[j] & 0xffff0000u);
            ds[j] = pack_bf16x2(px * (__uint_as_float(rp[2 * j]) - Di), py * (__uint_as_float(rp[2 * j + 1]) - Di));
          }
          *reinterpret_cast<uint4*>(addr) = make_uint4(ds[0], ds[1], ds[2], ds[3]);
        });
      }
      TR(7);
      tc::fence_before_sync();
      tc::fence_proxy_async();
      __syncthreads();
      TR(8);
      // ---- (4) dK_kt = dS^T Q_g ; (5) dQ_g = dS K
      if (warp == 0 && tc::elect_one()) {      // warp-uniform election: the MMAs issue without a per-lane replay loop
        tc::fence_after_sync();
        const uint32_t idesc_t = tc::make_idesc_bf16(128, DH, 1, 1);
        for (int t = 0; t < NQ; ++t)
          for (int kk = 0; kk < kq; ++kk) {
            const uint64_t a = tc::make_sdesc_sw128(sP + 2 * t * P_SLAB + kk * 2048, P_SLAB, 1024);
            const uint64_t b = tc::make_sdesc_sw128(sQ + g * 128 * ROWB + kk * 2048 + hoff, 8192, 1024);
            tc::mma_bf16_ss(tmem_base + COL_DK + t * 64, a, b, idesc_t, kk > 0);
          }
        const uint32_t idesc_q = tc::make_idesc_bf16(128, DH, 0, 1);
        for (int kk = 0; kk < NKP / 16; ++kk) {
          const uint64_t a = tc::make_sdesc_sw128(sP + (kk >> 2) * P_SLAB + (kk & 3) * 32, 16, 1024);
          const uint64_t b = tc::make_sdesc_sw128(sK + kk * 2048 + hoff, 8192, 1024);
          tc::mma_bf16_ss(tmem_base + COL_DQ, a, b, idesc_q, kk > 0);
        }
        tc::mma_commit(&bars[2]);
      }
      TR(9);
      tc::mbar_wait(&bars[2], phase);
      TR(10);
      if (tid == 0 && hd == HPB - 1 && g == NQ - 1 && item + (int)gridDim.x < n_items)
        issue_loads(item + gridDim.x);            // every MMA on this item's operands is done: refill during the drain
      tc::fence_after_sync();
      // ---- drain: dQ rows of this tile (columns split four ways), dK / dV partials into registers
      {
        constexpr int QW = DH / 4;                  // 8 or 16 dQ columns per thread
        uint32_t r[32];
        if constexpr (QW == 8) tc::tmem_ld_32x8(lane_base + COL_DQ + grp * QW, r);
        else tc::tmem_ld_32x16(lane_base + COL_DQ + grp * QW, r);
        tc::tmem_ld_wait();
        if (qok) {
          __nv_bfloat16* dq = dqkv + (size_t)(t0 + qrow) * ldq + h * DH + grp * QW;
#pragma unroll
          for (int j = 0; j < QW; j += 8) {
            uint4 o;
            o.x = pack_bf16x2(__uint_as_float(r[j]) * scale, __uint_as_float(r[j + 1]) * scale);
            o.y = pack_bf16x2(__uint_as_float(r[j + 2]) * scale, __uint_as_float(r[j + 3]) * scale);
            o.z = pack_bf16x2(__uint_as_float(r[j + 4]) * scale, __uint_as_float(r[j + 5]) * scale);
            o.w = pack_bf16x2(__uint_as_float(r[j + 6]) * scale, __uint_as_float(r[j + 7]) * scale);
            *reinterpret_cast<uint4*>(dq + j) = o;
          }
        }
        if (kt < NQ) {
          if constexpr (HC == 16) {
            tc::tmem_ld_32x16(lane_base + COL_DV + kt * DH + kh * HC, r);
            tc::tmem_ld_32x16(lane_base + COL_DK + kt * 64 + kh * HC, r + 16);
            tc::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) { accV[j] += __uint_as_float(r[j]); accK[j] += __uint_as_float(r[16 + j]); }
          } else {
            tc::tmem_ld_32x32(lane_base + COL_DV + kt * DH + kh * HC, r);
            tc::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) accV[j % HC] += (j < HC) ? __uint_as_float(r[j]) : 0.f;
            tc::tmem_ld_32x32(lane_base + COL_DK + kt * 64 + kh * HC, r);
            tc::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) accK[j % HC] += (j < HC) ? __uint_as_float(r[j]) : 0.f;
          }
        }
      }
      TR(11);
      tc::fence_before_sync();
      phase ^= 1;
      __syncthreads();                              // TMEM and smem free for the next tile / head / item
      TR(12);
    }
    // ---- store this thread's half key row
    const int krow = kt * 128 + row_in_tile;
    if (krow < S) {
      __nv_bfloat16* dk = dqkv + (size_t)(t0 + krow) * ldq + H * DH + h * DH + kh * HC;
      __nv_bfloat16* dv = dk + H * DH;
#pragma unroll
      for (int j = 0; j < HC; j += 8) {
        uint4 o;
        o.x = pack_bf16x2(accK[j] * scale, accK[j + 1] * scale); o.y = pack_bf16x2(accK[j + 2] * scale, accK[j + 3] * scale);
        o.z = pack_bf16x2(accK[j + 4] * scale, accK[j + 5] * scale); o.w = pack_bf16x2(accK[j + 6] * scale, accK[j + 7] * scale);
        *reinterpret_cast<uint4*>(dk + j) = o;
        o.x = pack_bf16x2(accV[j], accV[j + 1]); o.y = pack_bf16x2(accV[j + 2], accV[j + 3]);
        o.z = pack_bf16x2(accV[j + 4], accV[j + 5]); o.w = pack_bf16x2(accV[j + 6], accV[j + 7]);
        *reinterpret_cast<uint4*>(dv + j) = o;
      }
    }
    }  // heads of the box
  }
#ifdef EAVIT_TRACE
  if (blockIdx.x == 0 && tid == 0)
    for (int i = 0; i < 20; ++i) g_att_trace[i] += tr_acc[i];
#endif
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace eavit

using namespace eavit;

static int attention_tmaps(const void* qkv, const void* dout, long long T, int H, int Dh, CUtensorMap* tq, CUtensorMap* tdo) {
  int rc = make_tmap_bf16_2d(tq, qkv, (uint64_t)3 * H * Dh, (uint64_t)T, (uint64_t)3 * H * Dh * 2, BOX_ROWS);
  if (rc) return rc;
  if (dout != nullptr) rc = make_tmap_bf16_2d(tdo, dout, (uint64_t)H * Dh, (uint64_t)T, (uint64_t)H * Dh * 2, BOX_ROWS);
  return rc;
}

// Sequences per work item: > 1 only when every sequence has the same even length (total == nseq * max_len) and several fit
// into the `limit` tokens one item can hold.
static int att_pack(int nseq, int max_len, long long total_tokens, int limit) {
  if (total_tokens != (long long)nseq * max_len || (max_len & 1) || 2 * max_len > limit) return 1;
  return limit / max_len;
}

template <int DH, bool DROP, bool SEG>
static int launch_att_fwd2(const CUtensorMap& tq, const int* seq_start, int nseq, int H, float scale, void* out, float* lse,
                           const DropCfg& drop, int pack, int seg, cudaStream_t st) {
  static bool done = false;
  if (!done) { EAVIT_CUDA(cudaFuncSetAttribute(attention_fwd_tc_kernel<DH, DROP, SEG>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttSmem::TOTAL)); done = true; }
  const int items = ((nseq + pack - 1) / pack) * (H / (64 / DH));
  const int grid = items < kNumSMs ? items : kNumSMs;
  attention_fwd_tc_kernel<DH, DROP, SEG><<<grid, ATC_THREADS, AttSmem::TOTAL, st>>>(tq, seq_start, nseq, H, scale, (__nv_bfloat16*)out, lse, drop, pack, seg);
  EAVIT_LAUNCH_OK();
  return EAVIT_OK;
}
template <int DH, bool DROP>
static int launch_att_fwd(const CUtensorMap& tq, const int* seq_start, int nseq, int H, float scale, void* out, float* lse,
                          const DropCfg& drop, int pack, int seg, cudaStream_t st) {
  return pack > 1 ? launch_att_fwd2<DH, DROP, true>(tq, seq_start, nseq, H, scale, out, lse, drop, pack, seg, st)
                  : launch_att_fwd2<DH, DROP, false>(tq, seq_start, nseq, H, scale, out, lse, drop, 1, 0, st);
}
template <int DH, bool DROP, bool SEG>
static int launch_att_bwd2(const CUtensorMap& tq, const CUtensorMap& tdo, const float* lse, const int* seq_start, int nseq, int H,
                           float scale, void* dqkv, const DropCfg& drop, int pack, int seg, cudaStream_t st) {
  static bool done = false;
  if (!done) { EAVIT_CUDA(cudaFuncSetAttribute(attention_bwd_tc_kernel<DH, DROP, SEG>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttBwdSmem::TOTAL)); done = true; }
  const int items = ((nseq + pack - 1) / pack) * (H / (64 / DH));
  const int grid = items < kNumSMs ? items : kNumSMs;
  attention_bwd_tc_kernel<DH, DROP, SEG><<<grid, ATC_THREADS, AttBwdSmem::TOTAL, st>>>(tq, tdo, lse, seq_start, nseq, H, scale, (__nv_bfloat16*)dqkv, drop, pack, seg);
  EAVIT_LAUNCH_OK();
  return EAVIT_OK;
}
template <int DH, bool DROP>
static int launch_att_bwd(const CUtensorMap& tq, const CUtensorMap& tdo, const float* lse, const int* seq_start, int nseq, int H,
                          float scale, void* dqkv, const DropCfg& drop, int pack, int seg, cudaStream_t st) {
  return pack > 1 ? launch_att_bwd2<DH, DROP, true>(tq, tdo, lse, seq_start, nseq, H, scale, dqkv, drop, pack, seg, st)
                  : launch_att_bwd2<DH, DROP, false>(tq, tdo, lse, seq_start, nseq, H, scale, dqkv, drop, 1, 0, st);
}

// total tokens = seq_start[nseq]; the host mirror passes it so that no device read-back is needed
extern "C" int eavit_attention_fwd_tc(const void* qkv, const int* seq_start, int nseq, int max_len, long long total_tokens,
                                      int H, int Dh, float scale, void* out, float* lse, float drop_p,
                                      unsigned long long drop_seed, void* stream) {
  EAVIT_CHECK_ARG(qkv && seq_start && out && nseq > 0 && H > 0 && total_tokens > 0);
  EAVIT_CHECK_ARG(max_len > 0 && max_len <= ATC_MAXKEYS);
  EAVIT_CHECK_ARG((Dh == 32 && H % 2 == 0) || Dh == 64);
  EAVIT_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f);
  EAVIT_CHECK_ARG((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0);
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap tq, tdo;
  int rc = attention_tmaps(qkv, nullptr, total_tokens, H, Dh, &tq, &tdo);
  if (rc) return rc;
  const DropCfg drop = make_drop(drop_p, drop_seed);
  const int pack = att_pack(nseq, max_len, total_tokens, ATC_MAXKEYS), seg = max_len;
  if (Dh == 32) return drop.thresh ? launch_att_fwd<32, true>(tq, seq_start, nseq, H, scale, out, lse, drop, pack, seg, st)
                                   : launch_att_fwd<32, false>(tq, seq_start, nseq, H, scale, out, lse, drop, pack, seg, st);
  return drop.thresh ? launch_att_fwd<64, true>(tq, seq_start, nseq, H, scale, out, lse, drop, pack, seg, st)
                     : launch_att_fwd<64, false>(tq, seq_start, nseq, H, scale, out, lse, drop, pack, seg, st);
}

extern "C" int eavit_attention_bwd_tc(const void* qkv, const void* dout, const float* lse, const int* seq_start, int nseq,
                                      int max_len, long long total_tokens, int H, int Dh, float scale, void* dqkv, float drop_p,
                                      unsigned long long drop_seed, void* stream) {
  EAVIT_CHECK_ARG(qkv && dout && lse && seq_start && dqkv && nseq > 0 && H > 0 && total_tokens > 0);
  EAVIT_CHECK_ARG(max_len > 0 && max_len <= ATC_MAXKEYS);
  EAVIT_CHECK_ARG((Dh == 32 && H % 2 == 0) || (Dh == 64 && max_len <= 128));
  EAVIT_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f);
  EAVIT_CHECK_ARG((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(dout) & 15) == 0);
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap tq, tdo;
  int rc = attention_tmaps(qkv, dout, total_tokens, H, Dh, &tq, &tdo);
  if (rc) return rc;
  const DropCfg drop = make_drop(drop_p, drop_seed);
  const int pack = att_pack(nseq, max_len, total_tokens, Dh == 64 ? 128 : ATC_MAXKEYS), seg = max_len;
  if (Dh == 32) return drop.thresh ? launch_att_bwd<32, true>(tq, tdo, lse, seq_start, nseq, H, scale, dqkv, drop, pack, seg, st)
                                   : launch_att_bwd<32, false>(tq, tdo, lse, seq_start, nseq, H, scale, dqkv, drop, pack, seg, st);
  return drop.thresh ? launch_att_bwd<64, true>(tq, tdo, lse, seq_start, nseq, H, scale, dqkv, drop, pack, seg, st)
                     : launch_att_bwd<64, false>(tq, tdo, lse, seq_start, nseq, H, scale, dqkv, drop, pack, seg, st);
}

#ifdef EAVIT_TRACE
extern "C" int eavit_debug_att_trace(long long* host32, int reset) {
  if (reset) { long long z[32] = {0}; cudaMemcpyToSymbol(eavit::g_att_trace, z, sizeof(z)); return 0; }
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(host32, eavit::g_att_trace, 32 * sizeof(long long));
  return 0;
}
#endif
