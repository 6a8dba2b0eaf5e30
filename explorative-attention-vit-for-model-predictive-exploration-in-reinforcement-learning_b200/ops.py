"""Tensor-level wrappers over the C ABI (one Python function per ``eavit_*`` entry point).

All tensors must live on the current CUDA device; kernels are enqueued on the current torch stream.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import GemmArgs, check

U8, F32, F64, BF16 = 0, 1, 2, 3
ACT_NONE, ACT_GELU, ACT_GELU_BWD, ACT_LRELU, ACT_LRELU_BWD, ACT_RELU, ACT_RELU_BWD, ACT_MUL_AUX, ACT_GELU_SAVE_GRAD = range(9)
_DT = {torch.uint8: U8, torch.float32: F32, torch.float64: F64, torch.bfloat16: BF16}


# ------------------------------------------------------------------------------------------- launch profiler
# bench.py uses this to attribute device time to kernels with CUDA events on the launching stream.
_PROF = None


def profile_start():
    global _PROF
    _PROF = []


def profile_stop():
    """-> {label: (launches, total_ms, flops_per_launch)} ; synchronises."""
    global _PROF
    rec, _PROF = _PROF, None
    torch.cuda.synchronize()
    out = {}
    for label, e0, e1, flops in rec or []:
        n, ms, f = out.get(label, (0, 0.0, flops))
        out[label] = (n + 1, ms + e0.elapsed_time(e1), flops)
    return out


class _Timed:
    __slots__ = ("label", "flops", "e0")

    def __init__(self, label, flops=0.0):
        self.label, self.flops = label, flops

    def __enter__(self):
        if _PROF is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *a):
        if _PROF is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            _PROF.append((self.label, self.e0, e1, self.flops))


def _p(t: Optional[torch.Tensor]):
    if t is None:
        return None
    assert t.is_cuda, "eavit_b200 ops take CUDA tensors only (no CPU fallback)"
    return ctypes.c_void_p(t.data_ptr())


# The current stream's raw handle, once per launch (~150 launches per optimiser step): torch.cuda.current_stream() builds a
# Stream object and re-checks device availability on every call (14 us, 38 % of the host time of a step); the raw query is
# what torch itself uses underneath.
try:
    _raw_stream, _cur_dev = torch._C._cuda_getCurrentRawStream, torch._C._cuda_getDevice
    _raw_stream  # noqa: B018


    def _stream_handle() -> int:
        return _raw_stream(_cur_dev())
except AttributeError:                                     # pragma: no cover  (other torch builds)
    def _stream_handle() -> int:
        return torch.cuda.current_stream().cuda_stream


def _st():
    return ctypes.c_void_p(_stream_handle())


# ------------------------------------------------------------------------------------------- numerics
def gae_f64(reward: torch.Tensor, done: Optional[torch.Tensor], value: torch.Tensor, gamma: float, lam: float,
            kind: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """utils.py:42-67 with numpy promotion.  kind 0: reward f64 + done u8; kind 1: reward f32, no done."""
    E, T = reward.shape
    assert value.shape == (E, T + 1) and value.dtype == torch.float32 and value.is_contiguous()
    assert reward.is_contiguous() and reward.dtype == (torch.float64 if kind == 0 else torch.float32)
    if kind == 0:
        assert done is not None and done.dtype == torch.uint8 and done.shape == (E, T) and done.is_contiguous()
    ret = torch.empty(E * T, dtype=torch.float64, device=reward.device)
    adv = torch.empty_like(ret)
    check(_lib.lib().eavit_gae_f64(ctypes.c_int(kind), _p(reward), _p(done), _p(value), _p(ret), _p(adv), ctypes.c_int(E),
                                   ctypes.c_int(T), ctypes.c_double(gamma), ctypes.c_double(lam), _st()), "gae_f64")
    return ret, adv


def gae_f32(reward, done, value, gamma: float, lam: float):
    E, T = reward.shape
    assert reward.dtype == torch.float32 and value.dtype == torch.float32 and value.shape == (E, T + 1)
    assert reward.is_contiguous() and value.is_contiguous()
    ret = torch.empty(E * T, dtype=torch.float32, device=reward.device)
    adv = torch.empty_like(ret)
    check(_lib.lib().eavit_gae_f32(_p(reward), _p(done), _p(value), _p(ret), _p(adv), ctypes.c_int(E), ctypes.c_int(T),
                                   ctypes.c_float(gamma), ctypes.c_float(lam), _st()), "gae_f32")
    return ret, adv


def axpby_f64(a, b, ca: float, cb: float):
    out = torch.empty_like(a)
    check(_lib.lib().eavit_axpby_f64(_p(a), _p(b), _p(out), ctypes.c_longlong(a.numel()), ctypes.c_double(ca),
                                     ctypes.c_double(cb), _st()), "axpby_f64")
    return out


_ws_cache = {}


def _workspace(nbytes: int, device) -> torch.Tensor:
    key = (str(device),)
    w = _ws_cache.get(key)
    if w is None or w.numel() < nbytes:
        w = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _ws_cache[key] = w
    return w


def rms_update(x: torch.Tensor, mean: torch.Tensor, var: torch.Tensor, count: torch.Tensor):
    """utils.py:83-115 on device state (mean/var f64 [F], count f64 [1]); x [N, F] u8/f32/f64."""
    N, F = x.shape[0], x[0].numel()
    assert x.is_contiguous() and mean.numel() == F and var.numel() == F and count.numel() == 1
    ws = _workspace(int(_lib.lib().eavit_rms_workspace_bytes(N, F)), x.device)
    check(_lib.lib().eavit_rms_update(_p(x), ctypes.c_int(_DT[x.dtype]), ctypes.c_longlong(N), ctypes.c_int(F), _p(mean),
                                      _p(var), _p(count), _p(ws), _st()), "rms_update")


def rms_partial(x, shift):
    N, F = x.shape[0], x[0].numel()
    ws = _workspace(int(_lib.lib().eavit_rms_workspace_bytes(N, F)), x.device)
    s = torch.empty(F, dtype=torch.float64, device=x.device)
    q = torch.empty_like(s)
    check(_lib.lib().eavit_rms_partial(_p(x), ctypes.c_int(_DT[x.dtype]), ctypes.c_longlong(N), ctypes.c_int(F), _p(shift),
                                       _p(s), _p(q), _p(ws), _st()), "rms_partial")
    return s, q


def rms_merge(s, q, batch_count: float, mean, var, count):
    check(_lib.lib().eavit_rms_merge(_p(s), _p(q), ctypes.c_double(batch_count), ctypes.c_int(s.numel()), _p(mean), _p(var),
                                     _p(count), _st()), "rms_merge")


def obs_normalize(x, mean, var, out_dtype=torch.float32, out=None):
    """train.py:666/:855 ((x-mean)/sqrt(var)).clip(-5,5) -> float32 (agents.py:212/:298) or bf16."""
    N, F = x.shape[0], x[0].numel()
    assert x.is_contiguous()
    if out is None:
        out = torch.empty(x.shape, dtype=out_dtype, device=x.device)
    check(_lib.lib().eavit_obs_normalize(_p(x), ctypes.c_int(_DT[x.dtype]), ctypes.c_longlong(N), ctypes.c_int(F), _p(mean),
                                         _p(var), _p(out), ctypes.c_int(_DT[out.dtype]), _st()), "obs_normalize")
    return out


def reward_filter(int_reward, rewems, has_state: bool, gamma: float):
    """utils.py:118-128 + moments of train.py:738; returns moments f64[5] = mean, var, T, sum, sumsq."""
    E, T = int_reward.shape
    assert int_reward.dtype == torch.float32 and int_reward.is_contiguous() and rewems.numel() == E
    mom = torch.empty(5, dtype=torch.float64, device=int_reward.device)
    check(_lib.lib().eavit_reward_filter(_p(int_reward), _p(rewems), ctypes.c_int(int(has_state)), ctypes.c_int(E),
                                         ctypes.c_int(T), ctypes.c_float(gamma), _p(mom), None, _st()), "reward_filter")
    return mom


def scale_by_rsqrt_var(x, var):
    check(_lib.lib().eavit_scale_by_rsqrt_var(_p(x), ctypes.c_longlong(x.numel()), _p(var), _st()), "scale_by_rsqrt_var")
    return x


def intrinsic_mse(target, predict):
    N, R = target.shape
    assert target.dtype == torch.float32 and predict.dtype == torch.float32
    out = torch.empty(N, dtype=torch.float32, device=target.device)
    check(_lib.lib().eavit_intrinsic_mse(_p(target), _p(predict), _p(out), ctypes.c_int(N), ctypes.c_int(R), _st()),
          "intrinsic_mse")
    return out


# ------------------------------------------------------------------------------------------- GEMM
def gemm(A: torch.Tensor, B: torch.Tensor, *, a_mn: bool = False, b_mn: bool = False, bias=None, act: int = ACT_NONE,
         aux=None, residual=None, out_f32=None, out_bf16=None, out_pre=None, colsum=None, atomic: bool = False,
         split_k: int = 1, drop_p: float = 0.0, drop_seed: int = 0, ln=None):
    """C = epilogue(A . B^T) on tcgen05 (see include/eavit_b200.h: eavit_gemm_bf16).

    a_mn: A passed as the stored [K, M] matrix; b_mn: B passed as the stored [K, N] matrix.
    ln = (gamma, beta, mean_out, rstd_out, eps): fused LayerNorm of the residual row (N == 256): out_bf16 = LN(out_f32).
    Outputs must be preallocated by the caller ([M, N], row pitch = stride(0))."""
    assert A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16 and A.dim() == 2 and B.dim() == 2
    assert A.stride(1) == 1 and B.stride(1) == 1
    M, K = (A.shape[1], A.shape[0]) if a_mn else (A.shape[0], A.shape[1])
    N, Kb = (B.shape[1], B.shape[0]) if b_mn else (B.shape[0], B.shape[1])
    assert K == Kb, (A.shape, B.shape, a_mn, b_mn)
    ldc = None
    for t in (out_f32, out_bf16, out_pre, aux, residual):
        if t is not None:
            assert t.shape == (M, N) and t.stride(1) == 1, (t.shape, (M, N))
            ldc = t.stride(0) if ldc is None else ldc
            assert t.stride(0) == ldc
    g = GemmArgs(M=M, N=N, K=K, A=A.data_ptr(), lda=A.stride(0), a_mn=int(a_mn), B=B.data_ptr(), ldb=B.stride(0),
                 b_mn=int(b_mn), bias=None if bias is None else bias.data_ptr(),
                 aux_bf16=None if aux is None else aux.data_ptr(),
                 residual=None if residual is None else residual.data_ptr(),
                 out_f32=None if out_f32 is None else out_f32.data_ptr(),
                 out_bf16=None if out_bf16 is None else out_bf16.data_ptr(),
                 out_pre_bf16=None if out_pre is None else out_pre.data_ptr(),
                 colsum=None if colsum is None else colsum.data_ptr(),
                 ldc=ldc, drop_p=float(drop_p), drop_seed=int(drop_seed) & _M64, act=act, atomic_f32=int(atomic), split_k=split_k)
    if ln is not None:
        gamma, beta, mean_o, rstd_o, eps = ln
        g.ln_gamma, g.ln_beta = gamma.data_ptr(), beta.data_ptr()
        g.ln_mean = None if mean_o is None else mean_o.data_ptr()
        g.ln_rstd = None if rstd_o is None else rstd_o.data_ptr()
        g.ln_eps = float(eps)
    if _PROF is not None:
        with _Timed(f"gemm_bf16_tcgen05 M={M} N={N} K={K} {'mn' if a_mn else 'k'}{'mn' if b_mn else 'k'} act={act}"
                    f"{' splitk' if split_k > 1 else ''}{' +ln' if ln is not None else ''}", 2.0 * M * N * K):
            check(_lib.lib().eavit_gemm_bf16(ctypes.byref(g), _st()), "gemm_bf16")
        return
    check(_lib.lib().eavit_gemm_bf16(ctypes.byref(g), _st()), "gemm_bf16")


# ------------------------------------------------------------------------------------------- raw ABI access
# One-letter argument codes: p = device pointer (torch.Tensor or None), i = int, l = long long, f = float,
# d = double.  The trailing `void* stream` of every entry point is appended automatically.
_SPECS = {
    "eavit_layernorm_fwd": "plpppilppiif",
    "eavit_layernorm_bwd": "pilplppppl" "pl" "pl" "ppp" "fu" "ii",
    "eavit_dropout_apply": "plpiplii" "fu",
    "eavit_dropout_mask": "pliiii" "fu",
    "eavit_colsum": "pilpii",
    "eavit_gather_rows": "plpplii",
    "eavit_scatter_rows": "plppl" "pl" "ii",
    "eavit_scatter_add_rows": "plppl" "pl" "p" "ii" "fu",
    "eavit_cast_f32_bf16": "ppl",
    "eavit_add_f32": "pppl",
    "eavit_zero": "pl",
    "eavit_attention_fwd": "ppiiiifpp",
    "eavit_attention_fwd_tc": "ppiiliifpp" "fu",
    "eavit_attention_bwd": "pppppiiiifp",
    "eavit_attention_bwd_tc": "ppppiiliifp" "fu",
    "eavit_attention_bwd_tct": "pppppiiliifp" "fu",
    "eavit_attention_row0_fwd": "ppiiiifp" "fu",
    "eavit_attention_row0_bwd": "pppiiiifp" "fu",
    "eavit_patchify": "pipiiiiippfppp",
    "eavit_patchify_ln_bwd": "pipiiiiipppppp",
    "eavit_embed_assemble": "ppppiiiip",
    "eavit_embed_fused_fwd": "pipiiiii" "ppf" "pp" "ppf" "pp" "ppp" "ppp" "p" "i",
    "eavit_patch_ln_fold_bwd": "ppppp" "pppp" "ii",
    "eavit_embed_assemble_bwd": "piiiippppp",
    "eavit_embed_assemble_ln_bwd": "piiii" "pppp" "pppp" "ppp" "fu" "ppppppp",
    "eavit_sgemm_small": "pliplippplliiii".replace("ll", "l", 0),
    "eavit_heads_value_fwd": "ppppppiiip",
    "eavit_heads_value_bwd": "pppppiiipppppp",
    "eavit_combine_fwd": "ppiif",
    "eavit_combine_bwd": "ppiif",
    "eavit_ppo_loss": "ppppppppiifffpppp",
    "eavit_rnd_loss": "pppiifppp",
    "eavit_gather_batch": "piipppppppppp",
    "eavit_im2col": "pipiiiiiiipi",
    "eavit_im2col_nchw": "pipiiiiiiipi",
    "eavit_col2im_act": "ppiiiiiiipf",
    "eavit_flat_to_nhwc_act": "ppiiipf",
    "eavit_act_bwd_bf16": "ppfpl",
    "eavit_dropout_epoch_bump": "",
    "eavit_dropout_epoch_set": "i",
    "eavit_split3_rows": "pliipi",
    "eavit_bias_act_split3": "plpippii",
    "eavit_nhwc_to_flat_f32": "piiip",
    "eavit_col2im_lrelu": "ppiiiiiiip",
    "eavit_nhwc_to_flat": "piiip",
    "eavit_flat_to_nhwc_lrelu": "ppiiip",
    "eavit_adam_step": "ppppplpfffff",
    "eavit_adam_tick": "p",
    "eavit_adam_apply": "ppppplpfffff",
    "eavit_sumsq_f32": "plp",
    "eavit_clip_by_norm": "plpf",
}
_SPECS["eavit_sgemm_small"] = "pli" "pli" "pp" "pl" "iiiii"
_CT = {"p": ctypes.c_void_p, "i": ctypes.c_int, "l": ctypes.c_longlong, "f": ctypes.c_float, "d": ctypes.c_double,
       "u": ctypes.c_ulonglong}
_bound = {}


def _bind(name):
    fn = getattr(_lib.lib(), name)
    spec = _SPECS[name]
    fn.argtypes = [_CT[c] for c in spec] + [ctypes.c_void_p]
    fn.restype = ctypes.c_int
    _bound[name] = (fn, spec)
    return _bound[name]


_FLOPS_HINT = {}      # one-shot algorithmic FLOP count for the next profiled launch of an entry point (bench.py roofline)


def call(name: str, *args):
    """Invoke a C-ABI entry point with torch tensors / python scalars; raises on non-zero status."""
    fn, spec = _bound.get(name) or _bind(name)
    assert len(args) == len(spec), (name, len(args), len(spec))
    conv = []
    for a, c in zip(args, spec):
        if c == "p":
            if a is None:
                conv.append(None)
            else:
                assert a.is_cuda, f"{name}: CUDA tensors only"
                conv.append(a.data_ptr())
        else:
            conv.append(a)
    if _PROF is not None:
        with _Timed(name[len("eavit_"):], _FLOPS_HINT.pop(name, 0.0)):
            rc = fn(*conv, _stream_handle())
    else:
        rc = fn(*conv, _stream_handle())
    if rc != 0:
        check(rc, name[len("eavit_"):])


# ------------------------------------------------------------------------------------------- dropout
_M64 = (1 << 64) - 1


def site_seed(base: int, site: int) -> int:
    """Well-mixed 64-bit seed of one dropout site of one forward call (splitmix64 of base * 4096 + site)."""
    z = (int(base) * 4096 + int(site) + 0x9E3779B97F4A7C15) & _M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return z ^ (z >> 31)


def dropout_mask(n: int, ncols: int, p: float, seed: int, row0: int = 0, col0: int = 0, device="cuda") -> torch.Tensor:
    """The mask factors (0 or 1/(1-p)) the fused kernels apply at rows [row0, row0+n), columns [col0, col0+ncols)."""
    out = torch.empty(n, ncols, dtype=torch.float32, device=device)
    call("eavit_dropout_mask", out, ncols, row0, n, col0, ncols, float(p), int(seed) & _M64)
    return out


# ------------------------------------------------------------------------------------------- attention dispatch
def attention_fwd(qkv, seq_start, nseq, max_len, H, Dh, scale, out, lse, drop_p: float = 0.0, drop_seed: int = 0):
    """tcgen05 kernel when the sequence fits its TMEM plan (S <= 224), CUDA-core kernel for longer sequences."""
    if max_len <= 224:
        if _PROF is not None:      # 4 * S^2 * Dh per (sequence, head): QK^T and PV (S = max_len; cfg3: 196/197 vs 197)
            _FLOPS_HINT["eavit_attention_fwd_tc"] = 4.0 * nseq * H * max_len * max_len * Dh
        call("eavit_attention_fwd_tc", qkv, seq_start, nseq, max_len, qkv.shape[0], H, Dh, scale, out, lse, float(drop_p),
             int(drop_seed) & _M64)
    else:
        if drop_p > 0:
            raise NotImplementedError("attention-probability dropout is implemented by the tcgen05 kernels (sequence <= 224 tokens)")
        call("eavit_attention_fwd", qkv, seq_start, nseq, max_len, H, Dh, scale, out, lse)


def attention_bwd(qkv, out, dout, lse, seq_start, nseq, max_len, H, Dh, scale, dqkv, drop_p: float = 0.0, drop_seed: int = 0):
    if ((Dh == 32 and max_len <= 208 and H % 2 == 0) or (Dh == 64 and max_len <= 128)) and os.environ.get("EAVIT_ATTN_BWD", "t") == "t":
        if _PROF is not None:      # S, dP (recomputed), dV, dK, dQ: 5 contractions of 2 * S^2 * Dh
            _FLOPS_HINT["eavit_attention_bwd_tct"] = 10.0 * nseq * H * max_len * max_len * Dh
        call("eavit_attention_bwd_tct", qkv, out, dout, lse, seq_start, nseq, max_len, qkv.shape[0], H, Dh, scale, dqkv,
             float(drop_p), int(drop_seed) & _M64)
    elif (Dh == 32 and max_len <= 224) or (Dh == 64 and max_len <= 128):
        if _PROF is not None:
            _FLOPS_HINT["eavit_attention_bwd_tc"] = 10.0 * nseq * H * max_len * max_len * Dh
        call("eavit_attention_bwd_tc", qkv, dout, lse, seq_start, nseq, max_len, qkv.shape[0], H, Dh, scale, dqkv, float(drop_p),
             int(drop_seed) & _M64)
    else:
        if drop_p > 0:
            raise NotImplementedError("attention-probability dropout is implemented by the tcgen05 kernels")
        call("eavit_attention_bwd", qkv, out, dout, lse, seq_start, nseq, max_len, H, Dh, scale, dqkv)
