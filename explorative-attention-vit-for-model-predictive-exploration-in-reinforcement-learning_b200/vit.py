"""Drop-in for the reference's ``vit.py``: same class / enum names, constructor signature, parameter names
and init order (so ``state_dict`` keys, shapes and seeded initial values match), with the arithmetic done by
the sm_100a kernels.  Reference: vit.py:14-17 (ViT_Attn), :93-167 (ViT)."""
from __future__ import annotations

from enum import Enum

import torch
from torch import nn


class ViT_Attn(Enum):          # vit.py:14-17
    EXPLORATIVE_ATTN = 0
    EXPLOITATIVE_ATTN = 1
    CLS_ATTN = 2


def pair(t):
    return t if isinstance(t, tuple) else (t, t)


class _FeedForward(nn.Module):
    """Parameter holder mirroring vit.py:24-37 (indices 0, 1, 4 of ``net`` carry parameters)."""

    def __init__(self, dim, hidden_dim, dropout=0.0):
        super().__init__()
        self.net = nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, hidden_dim), nn.GELU(), nn.Dropout(dropout),
                                 nn.Linear(hidden_dim, dim), nn.Dropout(dropout))


class _Attention(nn.Module):
    """Parameter holder mirroring vit.py:39-58."""

    def __init__(self, dim, heads=8, dim_head=64, dropout=0.0):
        super().__init__()
        inner = dim_head * heads
        assert not (heads == 1 and dim_head == dim), "project_out=False variant is not used by any reference config"
        self.heads, self.scale = heads, dim_head ** -0.5
        self.norm = nn.LayerNorm(dim)
        self.to_qkv = nn.Linear(dim, inner * 3, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, dim), nn.Dropout(dropout))


class _Transformer(nn.Module):
    """Parameter holder mirroring vit.py:75-91."""

    def __init__(self, dim, depth, heads, dim_head, mlp_dim, dropout=0.0):
        super().__init__()
        self.norm = nn.LayerNorm(dim)
        self.layers = nn.ModuleList([])
        for _ in range(depth):
            self.layers.append(nn.ModuleList([_Attention(dim, heads=heads, dim_head=dim_head, dropout=dropout),
                                              _FeedForward(dim, mlp_dim, dropout=dropout)]))


class ViT(nn.Module):
    """ViT with exploration / exploitation tokens (reference vit.py:93-167).

    ``forward(img, attn_type)`` keeps the reference semantics including the token-prepend bug
    (vit.py:142/:146): EXPLORATIVE_ATTN -> 196 patch tokens, no token, no pos-emb, feature = patch 0;
    EXPLOITATIVE_ATTN -> ``exploration_token`` + pos-emb.  Both passes are computed together by the
    kernels; standalone calls return the requested half (inference only -- training goes through
    ``CnnActorCriticNetwork`` / ``RNDAgent`` which own the backward pass).
    """

    def __init__(self, *, image_size, patch_size, num_classes, dim, depth, heads, mlp_dim, pool="cls", channels=3,
                 dim_head=64, dropout=0.0, emb_dropout=0.0, use_explorativeAttn: bool = True):
        super().__init__()
        self.use_explorativeAttn = use_explorativeAttn
        ih, iw = pair(image_size)
        ph, pw = pair(patch_size)
        assert ih % ph == 0 and iw % pw == 0, "Image dimensions must be divisible by the patch size."
        assert ih == iw and ph == pw, "kernels assume square images / patches (84x84, reference configs)"
        assert pool == "cls" and num_classes == -1, "reference configs use pool='cls', num_classes=-1"
        num_patches = (ih // ph) * (iw // pw)
        patch_dim = channels * ph * pw
        self.to_patch_embedding = nn.Sequential(nn.Identity(), nn.LayerNorm(patch_dim), nn.Linear(patch_dim, dim),
                                                nn.LayerNorm(dim))
        self.pos_embedding = nn.Parameter(torch.randn(1, num_patches + 1, dim))
        if use_explorativeAttn:
            self.exploration_token = nn.Parameter(torch.randn(1, 1, dim))
            self.exploitation_token = nn.Parameter(torch.randn(1, 1, dim))
        else:
            self.cls_token = nn.Parameter(torch.randn(1, 1, dim))
        self.dropout = nn.Dropout(emb_dropout)
        self.transformer = _Transformer(dim, depth, heads, dim_head, mlp_dim, dropout)
        self.pool, self.num_classes = pool, num_classes
        self.hp = dict(image=ih, patch=ph, channels=channels, dim=dim, depth=depth, heads=heads, dim_head=dim_head,
                       mlp_dim=mlp_dim, dropout=dropout, emb_dropout=emb_dropout)
        self._rt = None

    def forward(self, img, attn_type: ViT_Attn):
        assert isinstance(attn_type, ViT_Attn), "attn_type must be of type ViT_Attn"
        from .model import standalone_vit_features
        if self.use_explorativeAttn:
            if attn_type == ViT_Attn.CLS_ATTN:
                raise Exception("explorative ViT takes EXPLORATIVE_ATTN / EXPLOITATIVE_ATTN")
        elif attn_type != ViT_Attn.CLS_ATTN:
            raise Exception("Must use attn_type=ViT_Attn.CLS_ATTN when self.use_explorativeAttn=True")   # vit.py:153
        f = standalone_vit_features(self, img)
        B = img.shape[0]
        return f[B:] if attn_type == ViT_Attn.EXPLOITATIVE_ATTN else f[:B]
