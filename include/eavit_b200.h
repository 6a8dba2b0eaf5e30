/*
 * eavit_b200 -- C ABI of the B200-native learner hot path.
 *
 * The reference (cangozpi/Explorative-Attention-ViT-...) is 100 % Python and has NO FFI / plugin
 * interface for this path: its boundary is the Python surface that train.py calls on
 * agents.py / model.py / vit.py / utils.py.  Each entry point below therefore cites the reference
 * Python call it replaces (file:line relative to the reference root); the host-side mirror of
 * that Python surface lives in the package next to csrc/ and is the only caller.
 *
 * Conventions
 *   - every function returns 0 on success, a negative EAVIT_E* code on failure, never throws;
 *   - all pointers are DEVICE pointers borrowed from the caller (PyTorch's caching allocator)
 *     unless the parameter name starts with "h_";
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*); no internal synchronisation;
 *   - the library keeps no global mutable state except cached function attributes.
 */
#ifndef EAVIT_B200_H
#define EAVIT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EAVIT_OK 0
#define EAVIT_EINVAL -1   /* bad argument (shape / alignment / enum) */
#define EAVIT_ECUDA -2    /* a CUDA runtime / driver call failed; see eavit_last_error() */
#define EAVIT_EUNSUPPORTED -3

/* element type tags for kernels that accept several input precisions */
#define EAVIT_U8 0
#define EAVIT_F32 1
#define EAVIT_F64 2
#define EAVIT_BF16 3

/* activation / epilogue tags of eavit_gemm_bf16 */
#define EAVIT_ACT_NONE 0
#define EAVIT_ACT_GELU 1        /* exact erf GELU (nn.GELU default, vit.py:30) */
#define EAVIT_ACT_GELU_BWD 2    /* v *= gelu'(aux) */
#define EAVIT_ACT_LRELU 3       /* LeakyReLU(0.01) (model.py:374) */
#define EAVIT_ACT_LRELU_BWD 4   /* v *= (aux > 0 ? 1 : 0.01) */
#define EAVIT_ACT_RELU 5
#define EAVIT_ACT_RELU_BWD 6    /* v *= (aux > 0) */
#define EAVIT_ACT_MUL_AUX 7     /* v *= aux          (backward of EAVIT_ACT_GELU_SAVE_GRAD) */
#define EAVIT_ACT_GELU_SAVE_GRAD 8   /* v = gelu(v) like EAVIT_ACT_GELU, but out_pre_bf16 receives gelu'(v) instead of v: the MLP
                                      * backward (vit.py:27-37) then multiplies instead of re-evaluating erf / exp */

const char* eavit_last_error(void);
int eavit_version(void);
/* number of kernels launched by this library since load (bench.py's gpu_launches claim) */
long long eavit_launch_count(void);

/* ------------------------------------------------------------------ numerics (utils.py, train.py) */

/* utils.py:42-67 make_train_data, UseGAE branch, with numpy's promotion rules reproduced so the
 * float64 outputs are bit-identical to the reference.
 *   kind 0 (extrinsic, train.py:748): reward f64 [E,T], done u8 [E,T], value f32 [E,T+1]
 *   kind 1 (intrinsic, train.py:757): reward f32 [E,T], done == zeros_like(reward) (ignored), value f32
 * ret / adv: f64 [E*T], flat index e*T+t (train.py:707-719). */
int eavit_gae_f64(int kind, const void* reward, const uint8_t* done, const float* value,
                  double* ret, double* adv, int E, int T, double gamma, double lam, void* stream);

/* Same recurrence in fp32 as a warp-shuffle affine scan (one warp per env); |err| <= 1e-5 rel. */
int eavit_gae_f32(const float* reward, const uint8_t* done /* may be NULL */, const float* value,
                  float* ret, float* adv, int E, int T, float gamma, float lam, void* stream);

/* train.py:767  out = a*ca + b*cb  (float64). */
int eavit_axpby_f64(const double* a, const double* b, double* out, long long n, double ca, double cb, void* stream);

/* utils.py:83-115 RunningMeanStd.update for the obs_rms: per-column population mean/var over the
 * N rows of x [N, F] (F = 84*84) followed by the Chan merge into (mean[F], var[F], count[1]), all
 * float64 device state.  x dtype = EAVIT_U8 / F32 / F64.  workspace: >= eavit_rms_workspace_bytes. */
long long eavit_rms_workspace_bytes(long long N, int F);
int eavit_rms_update(const void* x, int x_dtype, long long N, int F, double* mean, double* var,
                     double* count, void* workspace, void* stream);
/* Batch moments only (for the multi-GPU moment all-reduce): writes sum[F] and sumsq-about-shift[F]
 * where shift = current mean; merged by eavit_rms_merge after the all-reduce. */
int eavit_rms_partial(const void* x, int x_dtype, long long N, int F, const double* shift,
                      double* sum, double* sumsq, void* workspace, void* stream);
int eavit_rms_merge(const double* sum, const double* sumsq, double batch_count, int F,
                    double* mean, double* var, double* count, void* stream);

/* train.py:666 / :855  ((x - mean) / sqrt(var)).clip(-5, 5) evaluated in float64 like numpy, cast to
 * out_dtype (EAVIT_F32 as agents.py:212/:298, or EAVIT_BF16).  x [N,F] dtype U8/F32/F64. */
int eavit_obs_normalize(const void* x, int x_dtype, long long N, int F, const double* mean,
                        const double* var, void* out, int out_dtype, void* stream);

/* utils.py:118-128 + train.py:736-743: discounted forward filter over T per env (float32, state
 * rewems[E] carried across calls; has_state = 0 on the very first call), moments of the [T,E]
 * filter outputs in float64 -> moments[0]=mean, [1]=var(population), [2]=T (the count=T quirk),
 * [3]=sum, [4]=sum of squares (for the multi-GPU moment all-reduce).  moments has 5 doubles. */
int eavit_reward_filter(const float* int_reward /*[E,T]*/, float* rewems /*[E]*/, int has_state,
                        int E, int T, float gamma, double* moments /*[5]*/, void* workspace, void* stream);
/* train.py:743  total_int_reward /= sqrt(reward_rms.var)  (f32 / f64 -> f32, in place). */
int eavit_scale_by_rsqrt_var(float* x, long long n, const double* var /*[1]*/, void* stream);

/* agents.py:216  (target - predict).pow(2).mean(1) -> out[N] f32; inputs [N,R] fp32. */
int eavit_intrinsic_mse(const float* target, const float* predict, float* out, int N, int R, void* stream);

/* ------------------------------------------------------------------ dense tensor-core GEMM (tcgen05) */

/* C[M,N] = epilogue( A[M,K] . B[N,K]^T ), bf16 operands, fp32 accumulation in TMEM.
 *   a_mn = 0: A is row-major [M,K] (K contiguous, row pitch lda elements)
 *   a_mn = 1: A is given transposed, row-major [K,M] (M contiguous, pitch lda)   -- dW = dY^T X
 *   b_mn = 0: B is row-major [N,K] (pitch ldb)          -- nn.Linear weight, y = x W^T
 *   b_mn = 1: B is row-major [K,N] (pitch ldb)          -- dX = dY W
 * epilogue, per element v = acc:  v += bias[n];  out_pre_bf16 = v;  v = act(v, aux[m,n]);
 *   v *= dropmask(m,n) (drop_p > 0);  v += residual[m,n];  out_f32 / out_bf16 = v   (atomic_f32 != 0: out_f32 += v with red.add and
 *   split_k > 1 slices of K).  Any of bias/aux/residual/out_* may be NULL.  ldc = pitch of every
 *   [M,N] epilogue tensor in elements.  Requirements: pitches * 2 bytes % 16 == 0, N % 8 == 0.
 * Replaces every nn.Linear / matmul on the ViT path (vit.py:29-32, :52-57, :112) and the RND FCs. */
typedef struct {
  int M, N, K;
  const void* A; long long lda; int a_mn;
  const void* B; long long ldb; int b_mn;
  const float* bias;
  const void* aux_bf16;
  const float* residual;
  float* out_f32;
  void* out_bf16;
  void* out_pre_bf16;
  float* colsum;       /* optional [N]: colsum[n] += sum_m (final v) -- the bias gradient when this GEMM produces dY */
  long long ldc;
  float drop_p;        /* nn.Dropout probability applied to v after the activation (or its derivative) and before the */
  unsigned long long drop_seed; /* residual add; mask(m, n) is a pure function of (drop_seed, m, n) -- see eavit_dropout_mask */
  int act;
  int atomic_f32;
  int split_k;
  /* Fused PreNorm LayerNorm of the residual row (vit.py:28 / :47 after the residual adds of vit.py:88-89).  With ln_gamma != NULL
   * (N == 256, bias + residual + out_f32 + out_bf16 given, no activation): out_f32 = residual + dropout(A.B^T + bias) as usual and
   * out_bf16 = LayerNorm(out_f32 row) * ln_gamma + ln_beta, ln_mean / ln_rstd [M] (optional) = the row statistics the
   * LayerNorm backward needs.  One tile holds whole rows, so no standalone LayerNorm launch re-reads out_f32 from HBM. */
  const float* ln_gamma;
  const float* ln_beta;
  float* ln_mean;
  float* ln_rstd;
  float ln_eps;
} eavit_gemm_args;
int eavit_gemm_bf16(const eavit_gemm_args* args, void* stream);

/* ------------------------------------------------------------------ LayerNorm / small row ops */

/* nn.LayerNorm forward over rows of x fp32 [T,D] (vit.py:28,:47,:78,:113; HF layernorm_*).  y dtype
 * EAVIT_BF16 (GEMM operand) or EAVIT_F32; mean/rstd [T] saved for backward (both NULL to skip). */
int eavit_layernorm_fwd(const float* x, long long ldx, const float* gamma, const float* beta, void* y, int y_dtype,
                        long long ldy, float* mean, float* rstd, int T, int D, float eps, void* stream);
/* dx = dres + LN'(dy); dgamma/dbeta accumulated with atomics (+=).  dy dtype F32 or BF16; dres, dx,
 * dx_bf16, dgamma/dbeta optional.  dxsum (optional) += column sums of dx = the bias gradient of the Linear whose
 * output fed this residual stream (saves a separate pass over [T,D]).  drop_p > 0: dx_bf16 and dxsum carry the forward
 * dropout mask of that Linear's output (site seed drop_seed); the fp32 residual gradient dx does not. */
int eavit_layernorm_bwd(const void* dy, int dy_dtype, long long lddy, const float* x, long long ldx, const float* mean,
                        const float* rstd, const float* gamma, const float* dres, long long lddres, float* dx,
                        long long lddx, void* dx_bf16, long long lddxb, float* dgamma, float* dbeta, float* dxsum, float drop_p,
                        unsigned long long drop_seed, int T, int D, void* stream);
/* out[c] += sum_r x[r,c]  (bias gradients); x dtype BF16 or F32. */
int eavit_colsum(const void* x, int x_dtype, long long ldx, float* out, int T, int N, void* stream);
/* dst[i,:] = src[rows[i],:]  /  dst[rows[i],:] = src[i,:]  (pooled token x[:,0], vit.py:162). */
int eavit_gather_rows(const float* src, long long lds, const int* rows, float* dst, long long ldd, int n, int D, void* stream);
int eavit_scatter_rows(const float* src, long long lds, const int* rows, float* dst, long long ldd, void* dst_bf16,
                       long long lddb, int n, int D, void* stream);
/* dst[rows[i],:] += src[i,:] (rows unique) after a LayerNorm backward wrote dst: also refreshes the bf16 copy of those rows
 * (under the dropout mask (drop_p, drop_seed) of the linear layer below, vit.py:33) and adds the masked delta's column sums
 * to that layer's bias gradient.  The pooled-token residual gradient of the last layer (x[:, 0], vit.py:162). */
int eavit_scatter_add_rows(const float* src, long long lds, const int* rows, float* dst, long long ldd, void* dst_bf16 /* may be NULL */,
                           long long lddb, float* colsum /* may be NULL */, int n, int D, float drop_p, unsigned long long drop_seed,
                           void* stream);
int eavit_cast_f32_bf16(const float* in, void* out_bf16, long long n, void* stream);
int eavit_add_f32(const float* a, const float* b, float* out, long long n, void* stream);
int eavit_zero(void* ptr, long long bytes, void* stream);

/* ------------------------------------------------------------------ attention (vit.py:60-73) */

/* softmax(q k^T * scale) v per (sequence, head); qkv bf16 [T, 3*H*Dh] (q|k|v, each (h d)); out bf16
 * [T, H*Dh]; lse fp32 [T,H] (may be NULL for inference); seq_start int32 [nseq+1] token offsets;
 * max_len <= 256; Dh in {32, 64}. */
int eavit_attention_fwd(const void* qkv, const int* seq_start, int nseq, int max_len, int H, int Dh, float scale,
                        void* out, float* lse, void* stream);
/* Same contract on tcgen05 tensor cores (TMA-staged operands, S and O accumulators in TMEM, P staged as a bf16 MMA
 * operand); max_len <= 224.  total_tokens = T = seq_start[nseq] (the extent of the TMA tensor map: rows past T are
 * zero-filled, never read).  Dh = 32 needs an even H (two heads share one 128-byte staged row). */
int eavit_attention_fwd_tc(const void* qkv, const int* seq_start, int nseq, int max_len, long long total_tokens, int H, int Dh,
                           float scale, void* out, float* lse, float drop_p, unsigned long long drop_seed, void* stream);
int eavit_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse, const int* seq_start,
                        int nseq, int max_len, int H, int Dh, float scale, void* dqkv, void* stream);

/* Backward on tcgen05 (recomputes P from q, k, lse; does not need `out`).  Dh = 32: max_len <= 224; Dh = 64: max_len <= 128. */
int eavit_attention_bwd_tc(const void* qkv, const void* dout, const float* lse, const int* seq_start, int nseq, int max_len,
                           long long total_tokens, int H, int Dh, float scale, void* dqkv, float drop_p,
                           unsigned long long drop_seed, void* stream);
/* Backward of Attention.forward (vit.py:60-73 / HF ViTSelfAttention) with transposed scores (keys on the TMEM lanes; P~^T and dS^T are TMEM-resident A operands of dV / dK, one
 * softmax pass per key tile): Dh = 32 with max_len <= 208 or Dh = 64 with max_len <= 128; equal even-length sequences are
 * packed per work item under a block-diagonal mask.  Needs `out` (D = rowsum(out * dout), formed in-kernel). */
int eavit_attention_bwd_tct(const void* qkv, const void* out, const void* dout, const float* lse, const int* seq_start, int nseq,
                            int max_len, long long total_tokens, int H, int Dh, float scale, void* dqkv, float drop_p,
                            unsigned long long drop_seed, void* stream);
/* drop_p > 0 (both kernels): nn.Dropout on the attention probabilities (vit.py:45,70); mask element = (token row,
 * h * 256 + key) of the site seed, regenerated in the backward. */

/* Attention of the POOLED query only: the reference reads token 0 of every sequence after the last layer (x[:, 0],
 * vit.py:162), so that layer's attention output (and everything token-wise after it) is dead for query rows >= 1.
 *   fwd: out0 bf16 [nseq, H*Dh] = softmax(q_0 k^T * scale) v            (q_0 = first token of each sequence)
 *   bwd: dqkv bf16 [T, 3*H*Dh]  <- dout0 [nseq, H*Dh]; dq is zero except the first row of each sequence; dk, dv dense.
 * CUDA-core kernels (matrix-vector work), one warp per (sequence, head); max_len <= 512; drop_* as in the _tc kernels. */
int eavit_attention_row0_fwd(const void* qkv, const int* seq_start, int nseq, int max_len, int H, int Dh, float scale,
                             void* out0, float drop_p, unsigned long long drop_seed, void* stream);
int eavit_attention_row0_bwd(const void* qkv, const void* dout0, const int* seq_start, int nseq, int max_len, int H, int Dh,
                             float scale, void* dqkv, float drop_p, unsigned long long drop_seed, void* stream);

/* ------------------------------------------------------------------ dropout (vit.py:31,33,45,56,158) */

/* Every dropout site is a counter-based mask: keep(r, c) is a pure function of (seed, r, c) (16 hash bits per element,
 * p quantised to round(p*65536)/65536, kept values scaled by 1/(1-p)).  The fused kernels (GEMM epilogue, LayerNorm
 * backward, attention) evaluate it in place; these two entry points apply / materialise the SAME mask:
 *   eavit_dropout_apply: dst[i, c] = src[i, c] * mask(rows ? rows[i] : row0 + i, c)   (in place allowed)
 *   eavit_dropout_mask : out[i, j] = mask(row0 + i, col0 + j)  in {0, 1/(1-p)}          (tests / oracle hook) */
/* Dropout epoch: a per-device counter folded into every mask (mask = f(seed, epoch, row, col)); 0 unless set.  A CUDA
 * graph of the rollout forward (seeds frozen at capture) captures one `bump` so that every replay draws new masks, the way
 * each eager call of the reference's nn.Dropout does (vit.py:158; rollout in train mode, agents.py:187-195). */
int eavit_dropout_epoch_bump(void* stream);
int eavit_dropout_epoch_set(int value, void* stream);
int eavit_dropout_apply(const float* src, long long lds, const int* rows, int row0, float* dst, long long ldd, int n, int D,
                        float p, unsigned long long seed, void* stream);
int eavit_dropout_mask(float* out, long long ld, int row0, int n, int col0, int ncols, float p, unsigned long long seed,
                       void* stream);

/* ------------------------------------------------------------------ patch embedding (vit.py:109-158) */

/* [B,C,HW,HW] image (EAVIT_U8: value/255 as train.py:605; or EAVIT_F32) -> bf16 patches [B*np, C*P*P];
 * order_cpp = 0: lucidrains '(p1 p2 c)' (vit.py:110); 1: conv order '(c p1 p2)' (HF ViTPatchEmbeddings).
 * gamma/beta != NULL fuses LayerNorm(patch_dim) (vit.py:111) and saves mean/rstd [B*np].
 * sample_idx (int64 [B], may be NULL) gathers samples from a larger device-resident rollout. */
int eavit_patchify(const void* img, int img_dtype, const long long* sample_idx, int B, int C, int HW, int P, int order_cpp,
                   const float* gamma, const float* beta, float eps, void* out_bf16, float* mean, float* rstd,
                   void* stream);
/* dgamma/dbeta (+=) of that LayerNorm given dpln fp32 [B*np, PD]. */
int eavit_patchify_ln_bwd(const void* img, int img_dtype, const long long* sample_idx, int B, int C, int HW, int P,
                          int order_cpp, const float* gamma, const float* mean, const float* rstd, const float* dpln,
                          float* dgamma, float* dbeta, void* stream);
/* token prepend + positional add into the flat residual stream (mode 0: lucidrains explorative pair with
 * the reference's token bug, vit.py:141-156; 1: single CLS sequence; 2: HF pair, vit_hg.py:121-145). */
int eavit_embed_assemble(const float* e, const float* pos, const float* tokA, const float* tokB, int mode, int B, int np,
                         int D, float* x, void* stream);
/* The four steps above as ONE kernel for the lucidrains variants (mode 0 explorative pair / 1 CLS; patch_dim % 16 == 0,
 * patch_dim <= 192, dim == 256): vit.py:109-114 to_patch_embedding (Rearrange, LayerNorm(patch_dim), Linear, LayerNorm(dim)) +
 * vit.py:141-158 token prepend / positional add.  Image rows staged in shared memory, LayerNorm'ed patches written straight into
 * the swizzled A tile of a tcgen05 GEMM whose weight is TMA-loaded once per CTA, LayerNorm(dim) + assembly in the epilogue.
 * Also writes what the backward needs: pln bf16 [B*np, PD] (+ pmean, prstd), e0 fp32 [B*np, 256] (+ m3, r3).
 * pln_xhat != 0: pln receives the normalised patch xhat WITHOUT LayerNorm(patch_dim)'s affine (the GEMM still multiplies
 * g1 * xhat + b1): with G = dE^T xhat the Linear's weight gradient and that LayerNorm's dgamma / dbeta all follow from one
 * weight-gradient GEMM (eavit_patch_ln_fold_bwd) -- no dX GEMM and no second pass over the frames. */
int eavit_embed_fused_fwd(const void* img, int img_dtype, const long long* sample_idx /* may be NULL */, int B, int C, int HW, int P,
                          int mode, const float* g1, const float* b1, float eps1, const void* w_bf16, const float* bias,
                          const float* g3, const float* b3, float eps3, const float* pos, const float* tok, void* pln_bf16,
                          float* pmean, float* prstd, float* e0, float* m3, float* r3, float* x0, int pln_xhat, void* stream);
/* Backward of y = W (g1 * xhat + b1) + bias (vit.py:111-112: LayerNorm(patch_dim) then Linear) when the INPUT needs no gradient
 * (the frames): with G [N, K] = dY^T xhat (one split-K GEMM) and s [N] = column sums of dY,
 *   dW[j,k] += g1[k] G[j,k] + b1[k] s[j],   dbias[j] += s[j],   dg1[k] += sum_j W[j,k] G[j,k],   db1[k] += sum_j W[j,k] s[j]. */
int eavit_patch_ln_fold_bwd(const float* G, const float* s, const float* W, const float* g1, const float* b1, float* dW,
                            float* dbias, float* dg1, float* db1, int N, int K, void* stream);
/* Backward of eavit_embed_assemble (vit.py:141-158; vit_hg.py:121-145): g[b*np+n] = patch-token gradient summed over the
 * sequences (fp32 and / or bf16), dpos / dtokA / dtokB += the positional / token gradients summed over the samples.  One pass
 * over dx: a CTA per sequence position and sample chunk (D % 128 == 0, D <= 1024; other widths take two kernels). */
int eavit_embed_assemble_bwd(const float* dx, int mode, int B, int np, int D, float* g, void* g_bf16, float* dpos,
                             float* dtokA, float* dtokB, void* stream);
/* The same pass continued through the LayerNorm(dim) that ends to_patch_embedding (vit.py:113; D == 256): the patch-token
 * gradient stays in registers, de_bf16 [B*np, D] = LayerNorm'(g) (the dY operand of the patch Linear's dW / dX GEMMs),
 * dgamma / dbeta += that LayerNorm's parameter gradients, dbias (may be NULL) += column sums of de = the Linear's bias gradient
 * (vit.py:112).  e0 / mean / rstd: the LayerNorm's input and statistics as the forward stored them.  drop_p > 0: dx is read
 * under the embedding-dropout mask of the forward (vit.py:158; rows = flat token rows) instead of being masked in place first.
 * l1_dy_bf16 != NULL: the backward of the FIRST transformer layer's pre-attention LayerNorm (vit.py:47, layer 0) runs in front of
 * the pass: dx then is the residual gradient at that LayerNorm's input and the embedding-output gradient is formed per row as
 * dx + LN1'(l1_dy) (l1_dy = the QKV dX GEMM's bf16 output [T, D], l1_x = the embedding output, l1_mean / l1_rstd its
 * statistics); l1_dgamma / l1_dbeta += that LayerNorm's parameter gradients.  Replaces one eavit_layernorm_bwd over [T, D] and the
 * fp32 [T, D] round trip between the two kernels. */
int eavit_embed_assemble_ln_bwd(const float* dx, int mode, int B, int np, int D, const float* e0, const float* mean,
                                const float* rstd, const float* gamma, void* de_bf16, float* dgamma, float* dbeta, float* dbias,
                                float* dpos, float* dtokA, float* dtokB, float drop_p, unsigned long long drop_seed,
                                const void* l1_dy_bf16, const float* l1_x, const float* l1_mean, const float* l1_rstd,
                                const float* l1_gamma, float* l1_dgamma, float* l1_dbeta, void* stream);

/* ------------------------------------------------------------------ heads + losses (model.py, agents.py) */

/* small fp32 GEMM for the heads: C (+)= mask(relu?(op(A) op(B) + bias)). */
int eavit_sgemm_small(const float* A, long long lda, int transA, const float* B, long long ldb, int transB,
                      const float* bias, const float* mask_aux, float* C, long long ldc, int M, int N, int K, int relu,
                      int accumulate, void* stream);
/* v[r] = (E[r]+F[r]) . w(r) + b(r): critic(extra_layer(x)+x), model.py:276/:280 (w = A for r < split). */
int eavit_heads_value_fwd(const float* E, const float* F, const float* wA, const float* bA, const float* wB,
                          const float* bB, int split, int R, int D, float* v, void* stream);
int eavit_heads_value_bwd(const float* E, const float* F, const float* dv, const float* wA, const float* wB, int split,
                          int R, int D, float* dE, float* dF, float* dwA, float* dbA, float* dwB, float* dbB, void* stream);
int eavit_combine_fwd(const float* F, float* comb, int B, int D, float coef, void* stream);     /* model.py:284-288 */
int eavit_combine_bwd(const float* dcomb, float* dF, int B, int D, float coef, void* stream);
/* agents.py:455-493 fused forward+backward; stats fp32[16]: [1] actor [2] critic_ext [3] critic_int
 * [4] entropy [5] rnd [6] approx_kl [7] max_kl [8] clipfrac.  Gradients are multiplied by grad_scale. */
int eavit_ppo_loss(const float* logits, const float* old_logits, const long long* actions, const float* adv,
                   const float* v_ext, const float* v_int, const float* tgt_ext, const float* tgt_int, int B, int A,
                   float ppo_eps, float ent_coef, float grad_scale, float* dlogits, float* dv_ext, float* dv_int,
                   float* stats, void* stream);
/* agents.py:333-338 masked RND loss; stats[5] += loss; dpred bf16 [B,R]. */
int eavit_rnd_loss(const float* pred, const float* tgt, const float* mask, int B, int R, float grad_scale, void* dpred_bf16,
                   float* per_sample, float* stats, void* stream);

/* minibatch gather of the per-sample scalars (replaces agents.py:289-303 host indexing): out[i] = in[idx[i]]. */
int eavit_gather_batch(const long long* idx, int B, int A, const float* tgt_ext, const float* tgt_int, const float* adv,
                       const long long* actions, const float* old_logits, float* o_tgt_ext, float* o_tgt_int, float* o_adv,
                       long long* o_actions, float* o_old_logits, void* stream);

/* ------------------------------------------------------------------ RND convolutions (model.py:368-416) */
int eavit_im2col(const void* in, int in_dtype, const long long* sample_idx /* may be NULL */, int B, int H, int W, int C, int KH, int KW, int stride, void* col_bf16,
                 int split3 /* 1: write [hi|hi|lo] rows of 3*K (bf16x3 operand) */, void* stream);
/* bf16x3 split of fp32 rows: mode 0 [hi|hi|lo] (activations), mode 1 [hi|lo|hi] (weights); A3.B3^T = hi*hi+hi*lo+lo*hi.
 * Used for the RND towers, whose (Leaky)ReLU masks need near-fp32 pre-activations for gradient parity. */
int eavit_split3_rows(const float* in, long long ldi, int R, int K, void* out_bf16, int mode, void* stream);
/* Finisher of a split-K (atomically accumulated) fully-connected layer of the RND towers (model.py:380-416):
 * x <- act(x + bias) in place (fp32, act = EAVIT_ACT_NONE / RELU / LRELU); optional bf16 copy [R,K] and bf16x3 activation
 * rows [hi|hi|lo] [R,3K] for the next layer. */
int eavit_bias_act_split3(float* x, long long ldx, const float* bias, int act, void* out_bf16 /* may be NULL */,
                          void* out3_bf16 /* may be NULL */, int R, int K, void* stream);
int eavit_nhwc_to_flat_f32(const float* act, int B, int HW, int C, float* flat, void* stream);
int eavit_col2im_lrelu(const void* dcol_bf16, const void* act_bf16, int B, int H, int W, int C, int KH, int KW, int stride,
                       void* din_bf16, void* stream);
int eavit_nhwc_to_flat(const void* act_bf16, int B, int HW, int C, void* flat_bf16, void* stream);
/* ---- CNN actor-critic backbone (model.py:110-178, kept upstream as commented-out code; BASELINE configs[1]).  Same
 * im2col + tcgen05 GEMM scheme as the RND towers with a ReLU (slope 0) instead of LeakyReLU (slope 0.01):
 * eavit_im2col_nchw reads the NCHW frame stack [N,C,H,W] (uint8 raw frames, divided by 255 like np.float32(x)/255., or
 * float32) and writes the same [B*OH*OW, C*KH*KW] patch matrix as eavit_im2col; the *_act variants take the activation's
 * negative slope; eavit_act_bwd_bf16: out = bf16(dy * act'(y)) for the ReLU that ends the backbone (model.py:134). */
int eavit_im2col_nchw(const void* in, int in_dtype, const long long* sample_idx /* may be NULL */, int B, int H, int W, int C, int KH, int KW,
                      int stride, void* col_bf16, int split3, void* stream);
int eavit_col2im_act(const void* dcol_bf16, const void* act_bf16, int B, int H, int W, int C, int KH, int KW, int stride,
                     void* din_bf16, float slope, void* stream);
int eavit_flat_to_nhwc_act(const void* dflat_bf16, const void* act_bf16, int B, int HW, int C, void* dact_bf16, float slope, void* stream);
int eavit_act_bwd_bf16(const float* dy, const float* y, float slope, void* out_bf16, long long n, void* stream);
int eavit_flat_to_nhwc_lrelu(const void* dflat_bf16, const void* act_bf16, int B, int HW, int C, void* dact_bf16, void* stream);

/* ------------------------------------------------------------------ optimiser (agents.py:129,:508) */
int eavit_adam_step(float* p, const float* g, float* m, float* v, void* p_bf16, long long n, long long* step, float lr,
                    float beta1, float beta2, float eps, float grad_scale, void* stream);
/* Frozen tensors (train.py:261-263: requires_grad = False on model.feature.*; torch.optim.Adam skips tensors without a
 * gradient): eavit_adam_tick advances the shared step counter once, eavit_adam_apply updates one contiguous trainable
 * range [p, p+n) of the flat buffers without touching the counter.  eavit_adam_step == tick + apply over everything. */
int eavit_adam_tick(long long* step, void* stream);
int eavit_adam_apply(float* p, const float* g, float* m, float* v, void* p_bf16, long long n, const long long* step, float lr,
                     float beta1, float beta2, float eps, float grad_scale, void* stream);
int eavit_sumsq_f32(const float* x, long long n, float* out, void* stream);                       /* utils.py:141-170 */
int eavit_clip_by_norm(float* g, long long n, const float* sumsq, float max_norm, void* stream);  /* agents.py:497-499 */

#ifdef __cplusplus
}
#endif
#endif
