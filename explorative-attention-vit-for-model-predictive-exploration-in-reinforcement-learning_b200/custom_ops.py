"""``torch.ops.eavit_b200.*`` -- the hot-path kernels as torch custom ops (``torch.library``), with autograd.

The drop-in classes (``model.CnnActorCriticNetwork``, ``agents.RNDAgent``) drive the C ABI through ``engine`` because
they own flat parameter / gradient buffers; these ops expose the SAME kernels operator by operator for code that
composes them with ordinary torch autograd (e.g. a new head or a different block order on the reference's tensors):

    y      = torch.ops.eavit_b200.linear(x_bf16, w_bf16, bias_f32)                    # tcgen05 GEMM (vit.py:29,32,52,57)
    y, mean, rstd = torch.ops.eavit_b200.layer_norm(x_f32, gamma, beta, eps)         # fp32 stream -> bf16 operand
    o, lse = torch.ops.eavit_b200.attention(qkv_bf16, seq_start_i32, max_len, H, scale)
    ret, adv = torch.ops.eavit_b200.gae(reward_f32, value_f32, gamma, lam)           # utils.py:42-67 (fp32 scan)
    r      = torch.ops.eavit_b200.intrinsic_mse(target_f32, predict_f32)             # agents.py:216

Every op is registered for CUDA only: on a CPU tensor torch raises ``NotImplementedError`` -- there is no fallback.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor

from . import ops

_NS = "eavit_b200"


def _bf16_2d(t: Tensor) -> Tensor:
    assert t.dtype == torch.bfloat16 and t.dim() == 2
    return t if t.stride(1) == 1 else t.contiguous()


# ------------------------------------------------------------------------------------------- linear
@torch.library.custom_op(f"{_NS}::linear", mutates_args=(), device_types="cuda")
def linear(x: Tensor, weight: Tensor, bias: Optional[Tensor] = None) -> Tensor:
    x, weight = _bf16_2d(x), _bf16_2d(weight)
    y = torch.empty(x.shape[0], weight.shape[0], dtype=torch.bfloat16, device=x.device)
    ops.gemm(x, weight, bias=bias, out_bf16=y)
    return y


@linear.register_fake
def _(x, weight, bias=None):
    return x.new_empty(x.shape[0], weight.shape[0])


def _linear_setup(ctx, inputs, output):
    x, weight, bias = inputs
    ctx.save_for_backward(x, weight)
    ctx.has_bias = bias is not None


def _linear_bwd(ctx, dy):
    x, w = ctx.saved_tensors
    dy = _bf16_2d(dy)
    dx = torch.empty_like(x)
    ops.gemm(dy, w, b_mn=True, out_bf16=dx)                                     # dX = dY W
    dw32 = torch.zeros(w.shape, dtype=torch.float32, device=w.device)
    ops.gemm(dy, _bf16_2d(x), a_mn=True, b_mn=True, out_f32=dw32, atomic=True,
             split_k=max(1, min(148, (dy.shape[0] + 4095) // 4096)))            # dW = dY^T X (split-K)
    db = None
    if ctx.has_bias:
        db = torch.zeros(w.shape[0], dtype=torch.float32, device=w.device)
        ops.call("eavit_colsum", dy, ops.BF16, dy.stride(0), db, dy.shape[0], dy.shape[1])
    return dx, dw32.to(w.dtype), db


linear.register_autograd(_linear_bwd, setup_context=_linear_setup)


# ------------------------------------------------------------------------------------------- layer norm
@torch.library.custom_op(f"{_NS}::layer_norm", mutates_args=(), device_types="cuda")
def layer_norm(x: Tensor, gamma: Tensor, beta: Tensor, eps: float) -> Tuple[Tensor, Tensor, Tensor]:
    """nn.LayerNorm over the rows of the fp32 residual stream -> (bf16 GEMM operand, mean, rstd)."""
    assert x.dtype == torch.float32 and x.dim() == 2 and x.is_contiguous()
    T, D = x.shape
    y = torch.empty(T, D, dtype=torch.bfloat16, device=x.device)
    mean = torch.empty(T, dtype=torch.float32, device=x.device)
    rstd = torch.empty_like(mean)
    ops.call("eavit_layernorm_fwd", x, D, gamma, beta, y, ops.BF16, D, mean, rstd, T, D, float(eps))
    return y, mean, rstd


@layer_norm.register_fake
def _(x, gamma, beta, eps):
    return x.new_empty(x.shape, dtype=torch.bfloat16), x.new_empty(x.shape[0]), x.new_empty(x.shape[0])


def _ln_setup(ctx, inputs, output):
    x, gamma, beta, eps = inputs
    ctx.save_for_backward(x, gamma, output[1], output[2])


def _ln_bwd(ctx, dy, dmean, drstd):
    x, gamma, mean, rstd = ctx.saved_tensors
    T, D = x.shape
    dy = dy.contiguous()
    dt = ops.BF16 if dy.dtype == torch.bfloat16 else ops.F32
    dx = torch.empty_like(x)
    dg = torch.zeros(D, dtype=torch.float32, device=x.device)
    db = torch.zeros_like(dg)
    ops.call("eavit_layernorm_bwd", dy, dt, D, x, D, mean, rstd, gamma, None, D, dx, D, None, D, dg, db, None, 0.0, 0, T, D)
    return dx, dg, db, None


layer_norm.register_autograd(_ln_bwd, setup_context=_ln_setup)


# ------------------------------------------------------------------------------------------- attention
@torch.library.custom_op(f"{_NS}::attention", mutates_args=(), device_types="cuda")
def attention(qkv: Tensor, seq_start: Tensor, max_len: int, heads: int, scale: float) -> Tuple[Tensor, Tensor]:
    """softmax(q k^T * scale) v per (sequence, head); qkv bf16 [T, 3*H*Dh] -> (out bf16 [T, H*Dh], lse f32 [T, H])."""
    assert qkv.dtype == torch.bfloat16 and qkv.is_contiguous() and seq_start.dtype == torch.int32
    T = qkv.shape[0]
    Dh = qkv.shape[1] // (3 * heads)
    out = torch.empty(T, heads * Dh, dtype=torch.bfloat16, device=qkv.device)
    lse = torch.empty(T, heads, dtype=torch.float32, device=qkv.device)
    ops.attention_fwd(qkv, seq_start, seq_start.numel() - 1, max_len, heads, Dh, float(scale), out, lse)
    return out, lse


@attention.register_fake
def _(qkv, seq_start, max_len, heads, scale):
    T = qkv.shape[0]
    return qkv.new_empty(T, qkv.shape[1] // 3), qkv.new_empty(T, heads, dtype=torch.float32)


def _attn_setup(ctx, inputs, output):
    qkv, seq_start, max_len, heads, scale = inputs
    ctx.save_for_backward(qkv, seq_start, output[0], output[1])
    ctx.cfg = (max_len, heads, scale)


def _attn_bwd(ctx, dout, dlse):
    qkv, seq_start, out, lse = ctx.saved_tensors
    max_len, heads, scale = ctx.cfg
    Dh = qkv.shape[1] // (3 * heads)
    dqkv = torch.empty_like(qkv)
    ops.attention_bwd(qkv, out, dout.contiguous().to(torch.bfloat16), lse, seq_start, seq_start.numel() - 1, max_len, heads, Dh,
                      float(scale), dqkv)
    return dqkv, None, None, None, None


attention.register_autograd(_attn_bwd, setup_context=_attn_setup)


# ------------------------------------------------------------------------------------------- rollout numerics
@torch.library.custom_op(f"{_NS}::gae", mutates_args=(), device_types="cuda")
def gae(reward: Tensor, value: Tensor, gamma: float, lam: float) -> Tuple[Tensor, Tensor]:
    """utils.py:42-67 as an fp32 warp-shuffle scan over T, parallel over envs (no done mask: the intrinsic stream)."""
    return ops.gae_f32(reward.contiguous(), None, value.contiguous(), float(gamma), float(lam))


@gae.register_fake
def _(reward, value, gamma, lam):
    n = reward.shape[0] * reward.shape[1]
    return reward.new_empty(n), reward.new_empty(n)


@torch.library.custom_op(f"{_NS}::intrinsic_mse", mutates_args=(), device_types="cuda")
def intrinsic_mse(target: Tensor, predict: Tensor) -> Tensor:
    """agents.py:216  (target - predict).pow(2).mean(1)."""
    return ops.intrinsic_mse(target.contiguous(), predict.contiguous())


@intrinsic_mse.register_fake
def _(target, predict):
    return target.new_empty(target.shape[0])


@torch.library.custom_op(f"{_NS}::obs_normalize", mutates_args=(), device_types="cuda")
def obs_normalize(x: Tensor, mean: Tensor, var: Tensor) -> Tensor:
    """train.py:666/:855  ((x - mean) / sqrt(var)).clip(-5, 5) in float64 arithmetic -> float32."""
    return ops.obs_normalize(x.contiguous(), mean, var)


@obs_normalize.register_fake
def _(x, mean, var):
    return x.new_empty(x.shape, dtype=torch.float32)
