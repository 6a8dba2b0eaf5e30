"""Drop-in for the reference's ``vit_hg.py`` (HuggingFace-style ViT with explorative attention).

Parameter tree and init follow vit_hg.py:46-66 (embeddings), :179-224 (_init_weights: trunc-normal
std = initializer_range for Linear / Conv / tokens / position embeddings, LayerNorm = (1, 0)) and
:236-255 (ViT_ExplorativeAttn: embeddings, encoder, layernorm, pooler).  The encoder arithmetic that
the reference imports from ``transformers.models.vit.modeling_vit`` (pinned 4.37.0) is implemented by
the same sm_100a kernels as the lucidrains variant (fused q|k|v GEMM with bias, eps = layer_norm_eps).
``transformers`` is NOT imported here: ``config`` only needs the ViTConfig attributes read below.
"""
from __future__ import annotations

from types import SimpleNamespace

import torch
from torch import nn


def ViTConfigLike(**kw):
    """Minimal stand-in for ``transformers.ViTConfig`` (a real ViTConfig instance works too)."""
    d = dict(hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072, hidden_act="gelu",
             hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, initializer_range=0.02, layer_norm_eps=1e-12,
             image_size=224, patch_size=16, num_channels=3, qkv_bias=True, encoder_stride=16)
    d.update(kw)
    return SimpleNamespace(**d)


class _PatchEmbeddings(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.projection = nn.Conv2d(c.num_channels, c.hidden_size, kernel_size=c.patch_size, stride=c.patch_size)
        self.num_patches = (c.image_size // c.patch_size) ** 2


class ViTEmbeddings_ExplorativeAttn(nn.Module):
    """vit_hg.py:46-66."""

    def __init__(self, config, use_mask_token: bool = False):
        super().__init__()
        assert not use_mask_token
        self.exploration_token = nn.Parameter(torch.randn(1, 1, config.hidden_size))
        self.exploitation_token = nn.Parameter(torch.randn(1, 1, config.hidden_size))
        self.patch_embeddings = _PatchEmbeddings(config)
        self.position_embeddings = nn.Parameter(torch.randn(1, self.patch_embeddings.num_patches + 1, config.hidden_size))
        self.config = config


class _SelfAttention(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.query = nn.Linear(c.hidden_size, c.hidden_size, bias=c.qkv_bias)
        self.key = nn.Linear(c.hidden_size, c.hidden_size, bias=c.qkv_bias)
        self.value = nn.Linear(c.hidden_size, c.hidden_size, bias=c.qkv_bias)


class _Dense(nn.Module):
    def __init__(self, i, o):
        super().__init__()
        self.dense = nn.Linear(i, o)


class _Attention(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.attention = _SelfAttention(c)
        self.output = _Dense(c.hidden_size, c.hidden_size)


class _Layer(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.attention = _Attention(c)
        self.intermediate = _Dense(c.hidden_size, c.intermediate_size)
        self.output = _Dense(c.intermediate_size, c.hidden_size)
        self.layernorm_before = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)
        self.layernorm_after = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)


class _Encoder(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.layer = nn.ModuleList([_Layer(c) for _ in range(c.num_hidden_layers)])


class ViT_ExplorativeAttn(nn.Module):
    """vit_hg.py:227-374.  ``forward(pixel_values)`` returns two objects whose ``[0]`` is the final-LayerNorm'ed
    sequence feature ``[B, 1, D]`` restricted to token 0 -- callers use ``out[0][:, 0, :]`` (model.py:316,:320)."""

    def __init__(self, config, add_pooling_layer: bool = True, use_mask_token: bool = False, use_explorativeAttn: bool = True):
        super().__init__()
        assert use_explorativeAttn, "only the explorative HF variant is wired (no reference config uses HF + CLS)"
        assert config.qkv_bias, "fused q|k|v GEMM expects biases (ViTHG_qkv_bias = True in every reference config)"
        self.use_explorativeAttn = use_explorativeAttn
        self.config = config
        self.embeddings = ViTEmbeddings_ExplorativeAttn(config, use_mask_token=use_mask_token)
        self.encoder = _Encoder(config)
        self.layernorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.pooler = _Dense(config.hidden_size, config.hidden_size) if add_pooling_layer else None
        self.apply(self._init_weights)
        self.hp = dict(image=config.image_size, patch=config.patch_size, channels=config.num_channels, dim=config.hidden_size,
                       depth=config.num_hidden_layers, heads=config.num_attention_heads,
                       dim_head=config.hidden_size // config.num_attention_heads, mlp_dim=config.intermediate_size,
                       dropout=config.hidden_dropout_prob, emb_dropout=config.hidden_dropout_prob,
                       attn_dropout=config.attention_probs_dropout_prob, act_dropout=0.0)   # HF has no dropout after GELU
        self._rt = None

    def _init_weights(self, module):   # vit_hg.py:179-224
        std = self.config.initializer_range
        if isinstance(module, (nn.Linear, nn.Conv2d)):
            nn.init.trunc_normal_(module.weight.data, mean=0.0, std=std)
            if module.bias is not None:
                module.bias.data.zero_()
        elif isinstance(module, nn.LayerNorm):
            module.bias.data.zero_()
            module.weight.data.fill_(1.0)
        elif isinstance(module, ViTEmbeddings_ExplorativeAttn):
            for t in (module.position_embeddings, module.exploration_token, module.exploitation_token):
                nn.init.trunc_normal_(t.data, mean=0.0, std=std)

    def forward(self, pixel_values):
        from .model import standalone_vit_features
        f = standalone_vit_features(self, pixel_values)
        B = pixel_values.shape[0]
        return (f[:B].unsqueeze(1),), (f[B:].unsqueeze(1),)
