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
 *   v += residual[m,n];  out_f32 / out_bf16 = v   (atomic_f32 != 0: out_f32 += v with red.add and
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
  long long ldc;
  int act;
  int atomic_f32;
  int split_k;
} eavit_gemm_args;
int eavit_gemm_bf16(const eavit_gemm_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif
