"""Forward / backward orchestration of the learner hot path on top of the C-ABI kernels.

Everything numerical happens in ``libeavit_b200.so``; this file only sequences launches and owns the
device buffers:

  * ``ParamStore``  -- ONE flat fp32 buffer for all trainable tensors (+ bf16 shadow for the tensor-core
    GEMMs, + flat gradient, + Adam moments).  The optimiser is one launch and the multi-GPU gradient
    exchange is one NCCL all-reduce over ``store.grad`` (replaces the reference's never-armed DDP reducer,
    SURVEY fact 5).
  * ``ViTEncoder``  -- patch-embed -> token/pos assembly -> depth x (LN, QKV GEMM, attention, out-proj +
    residual, LN, MLP1 + GELU, MLP2 + residual) -> pooled-token LN.  The explorative (S = np) and
    exploitative (S = np+1) passes of model.py:275/:279 share weights, so both run as ONE flat token
    batch [B*np + B*(np+1), D]: every GEMM sees both passes in one launch and the patch embedding is
    computed once instead of twice.
  * ``Heads`` / ``RNDNet`` -- PPO heads (fp32) and the RND conv towers (im2col + tcgen05 GEMM).

Reference call sites: model.py:266-354, vit.py:136-167, vit_hg.py:277-374, agents.py:333-508.
"""
from __future__ import annotations

import os
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import ops
from .config import HotPathConfig
from .ops import call

F32, BF16 = ops.F32, ops.BF16


def _pad8(n: int) -> int:
    return (n + 7) // 8 * 8


class ParamStore:
    """Flat storage for a set of named tensors (fp32 master, bf16 shadow, grad, Adam m / v)."""

    def __init__(self, shapes: "OrderedDict[str, Tuple[int, ...]]", device, trainable: bool = True):
        self.shapes = OrderedDict(shapes)
        self.offsets: Dict[str, int] = {}
        off = 0
        for name, shp in self.shapes.items():
            self.offsets[name] = off
            n = 1
            for s in shp:
                n *= s
            off += _pad8(n)
        self.numel = max(off, 8)
        self.device = device
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=device)
        self.flat_bf16 = torch.zeros(self.numel, dtype=torch.bfloat16, device=device)
        self.trainable = trainable
        if trainable:
            self.grad = torch.zeros(self.numel, dtype=torch.float32, device=device)
            self.m = torch.zeros(self.numel, dtype=torch.float32, device=device)
            self.v = torch.zeros(self.numel, dtype=torch.float32, device=device)
            self.step = torch.zeros(1, dtype=torch.int64, device=device)
        self._views = {n: self._view(self.flat, n) for n in self.shapes}
        self._bviews = {n: self._view(self.flat_bf16, n) for n in self.shapes}
        self._gviews = {n: self._view(self.grad, n) for n in self.shapes} if trainable else {}

    def _view(self, buf, name):
        shp = self.shapes[name]
        n = 1
        for s in shp:
            n *= s
        o = self.offsets[name]
        return buf[o:o + n].view(shp)

    def w(self, name) -> torch.Tensor:        # fp32 master
        return self._views[name]

    def b16(self, name) -> torch.Tensor:      # bf16 shadow
        return self._bviews[name]

    def g(self, name) -> torch.Tensor:        # fp32 gradient
        return self._gviews[name]

    def span(self, names: Sequence[str], buf: str, shape) -> torch.Tensor:
        """View over several ADJACENT tensors (e.g. HF query|key|value weights as one [3D, D] matrix)."""
        o = self.offsets[names[0]]
        n = 1
        for s in shape:
            n *= s
        exp = o
        for nm in names:
            assert self.offsets[nm] == exp, "tensors are not adjacent in the flat store"
            k = 1
            for s in self.shapes[nm]:
                k *= s
            assert k % 8 == 0
            exp += k
        assert exp - o == n
        return getattr(self, buf)[o:o + n].view(shape)

    def sync_shadow(self):
        call("eavit_cast_f32_bf16", self.flat, self.flat_bf16, self.numel)

    def zero_grad(self):
        call("eavit_zero", self.grad, self.numel * 4)

    def adam_step(self, lr: float, grad_scale: float = 1.0, beta1=0.9, beta2=0.999, eps=1e-8, ranges=None):
        """One fused Adam launch over the whole flat buffer, or -- with frozen tensors -- one step-counter tick plus one
        launch per contiguous trainable range (``trainable_ranges``): frozen tensors keep weights AND moments, like
        torch.optim.Adam skipping parameters without a gradient (train.py:261-263)."""
        if ranges is None:
            call("eavit_adam_step", self.flat, self.grad, self.m, self.v, self.flat_bf16, self.numel, self.step,
                 lr, beta1, beta2, eps, grad_scale)
            return
        call("eavit_adam_tick", self.step)
        for lo, hi in ranges:
            call("eavit_adam_apply", self.flat[lo:hi], self.grad[lo:hi], self.m[lo:hi], self.v[lo:hi], self.flat_bf16[lo:hi],
                 hi - lo, self.step, lr, beta1, beta2, eps, grad_scale)

    def name_ranges(self, names) -> List[Tuple[int, int]]:
        """Merged [lo, hi) element ranges (padding included) covered by ``names``, in offset order."""
        order = sorted(self.offsets.items(), key=lambda kv: kv[1])
        ends = [o for _, o in order[1:]] + [self.numel]
        pick = set(names)
        out: List[Tuple[int, int]] = []
        for (n, lo), hi in zip(order, ends):
            if n not in pick:
                continue
            if out and out[-1][1] == lo:
                out[-1] = (out[-1][0], hi)
            else:
                out.append((lo, hi))
        return out

    def adopt_optimizer_state(self, old: "ParamStore") -> bool:
        """Carry Adam moments and the step counter over from the store this one replaces (a Runtime rebuild after
        ``.to()`` / ``load_state_dict(assign=True)``).  Returns False when the layouts differ."""
        if not (self.trainable and old.trainable) or list(self.shapes.items()) != list(old.shapes.items()):
            return False
        self.m.copy_(old.m)
        self.v.copy_(old.v)
        self.step.copy_(old.step)
        return True


class _Buffers:
    """Named, shape-checked scratch tensors, allocated once per (batch) configuration."""

    def __init__(self, device):
        self.device = device
        self.t: Dict[str, torch.Tensor] = {}

    def get(self, name, shape, dtype) -> torch.Tensor:
        t = self.t.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = torch.empty(shape, dtype=dtype, device=self.device)
            self.t[name] = t
        return t


def _split_k(M: int, N: int, K: int) -> int:
    """K splits of a weight-gradient GEMM: the smallest count whose work items (tiles * splits) fill the persistent
    148-CTA grid to >= 90 % in whole waves (149 items would cost a second wave; 96 tiles alone leave a third of the SMs
    idle, 96 * 3 = 288 items fill two waves to 97 %), with at least 8 k-blocks per split."""
    mt = (M + 127) // 128
    if N > 128 and mt >= 2 and mt * ((N + 255) // 256) >= 4 and not os.environ.get("EAVIT_NO_PAIR"):
        mt = (mt + 1) // 2           # the 256-wide kernel takes 128-row tiles in pairs (both TMEM accumulators, B read once)
    tiles = mt * ((N + 255) // 256 if N > 128 else 1)
    kb = (K + 63) // 64
    one_wave = max(1, min(148 // tiles, kb))
    if tiles * one_wave >= 0.85 * 148:
        return one_wave
    best, best_eff = one_wave, tiles * one_wave / 148.0
    for sp in range(one_wave + 1, 149):
        if kb // sp < 8:
            break
        items = tiles * sp
        eff = items / (148.0 * ((items + 147) // 148))
        if eff >= 0.9:
            return sp
        if eff > best_eff + 1e-9:
            best, best_eff = sp, eff
    return best


def linear_fwd(x16, w16, *, bias=None, act=ops.ACT_NONE, residual=None, out_f32=None, out_bf16=None, out_pre=None,
               drop_p=0.0, drop_seed=0, ln=None):
    """y = dropout(act(x W^T + b)) (+ residual) on the tcgen05 GEMM; ``ln``: the LayerNorm that consumes y, fused."""
    ops.gemm(x16, w16, bias=bias, act=act, residual=residual, out_f32=out_f32, out_bf16=out_bf16, out_pre=out_pre,
             drop_p=drop_p, drop_seed=drop_seed, ln=ln)


def linear_bwd(dy16, x16, w16, *, dW, db=None, dx_f32=None, dx_bf16=None, act=ops.ACT_NONE, aux=None, dx_colsum=None,
               drop_p=0.0, drop_seed=0):
    """dW += dy^T x (split-K, MN-major operands), db += colsum(dy), dx = act'(dy W); dx_colsum += colsum(dx) fused in
    the dX epilogue (= the bias gradient of the layer below, whose output gradient dx is)."""
    M, K = dW.shape
    ops.gemm(dy16, x16, a_mn=True, b_mn=True, out_f32=dW, atomic=True, split_k=_split_k(M, K, dy16.shape[0]))
    if db is not None:
        call("eavit_colsum", dy16, BF16, dy16.stride(0), db, dy16.shape[0], dy16.shape[1])
    if dx_f32 is not None or dx_bf16 is not None:
        ops.gemm(dy16, w16, b_mn=True, act=act, aux=aux, out_f32=dx_f32, out_bf16=dx_bf16, colsum=dx_colsum,
                 drop_p=drop_p, drop_seed=drop_seed)


class ViTEncoder:
    """Both ViT variants of the reference behind one flat-token implementation."""

    def __init__(self, cfg: HotPathConfig, store: ParamStore, prefix: str = "model.feature."):
        self.cfg, self.store, self.pre = cfg, store, prefix
        self.dev = store.device
        self.buf: Dict[int, _Buffers] = {}
        c = cfg
        self.inner = c.heads * c.dim_head
        if c.impl == "lucidrains":
            self.mode = 0 if c.use_explorative else 1
        else:
            assert c.use_explorative, "HF variant is only wired for use_explorativeAttn=True (no shipped config uses CLS)"
            self.mode = 2
        self.nseq_per_sample = 1 if self.mode == 1 else 2
        # Only token 0 of each sequence is read after the last layer (x[:, 0], vit.py:162): that layer's out-projection,
        # LayerNorm and MLP run on the pooled rows only and its attention with one query row per (sequence, head) --
        # identical loss and gradients, a third of the ViT's token-wise work and one of three attention launches removed.
        self.prune_last = os.environ.get("EAVIT_PRUNE_LAST", "1") != "0"
        self.fuse_embed_bwd = os.environ.get("EAVIT_FUSE_EMBED_BWD", "1") != "0"
        self.fuse_embed_ln1 = os.environ.get("EAVIT_FUSE_EMBED_LN1", "1") != "0"
        # layer parameter names
        p = prefix
        self.L = []
        for i in range(c.depth):
            if c.impl == "lucidrains":
                a, m = p + f"transformer.layers.{i}.0.", p + f"transformer.layers.{i}.1.net."
                self.L.append(dict(ln1=(a + "norm.weight", a + "norm.bias"), qkv_w=a + "to_qkv.weight", qkv_b=None,
                                   o_w=a + "to_out.0.weight", o_b=a + "to_out.0.bias", ln2=(m + "0.weight", m + "0.bias"),
                                   w1=m + "1.weight", b1=m + "1.bias", w2=m + "4.weight", b2=m + "4.bias"))
            else:
                l = p + f"encoder.layer.{i}."
                self.L.append(dict(ln1=(l + "layernorm_before.weight", l + "layernorm_before.bias"),
                                   qkv_w=[l + f"attention.attention.{n}.weight" for n in ("query", "key", "value")],
                                   qkv_b=[l + f"attention.attention.{n}.bias" for n in ("query", "key", "value")],
                                   o_w=l + "attention.output.dense.weight", o_b=l + "attention.output.dense.bias",
                                   ln2=(l + "layernorm_after.weight", l + "layernorm_after.bias"),
                                   w1=l + "intermediate.dense.weight", b1=l + "intermediate.dense.bias",
                                   w2=l + "output.dense.weight", b2=l + "output.dense.bias"))

    # ---- parameter views -----------------------------------------------------------------------
    def _qkv(self, L, kind):
        s, D, I = self.store, self.cfg.dim, self.inner
        if isinstance(L["qkv_w"], str):
            return {"w16": s.b16(L["qkv_w"]), "gw": s.g(L["qkv_w"]), "b": None, "gb": None}[kind]
        if kind == "w16":
            return s.span(L["qkv_w"], "flat_bf16", (3 * I, D))
        if kind == "gw":
            return s.span(L["qkv_w"], "grad", (3 * I, D))
        if kind == "b":
            return s.span(L["qkv_b"], "flat", (3 * I,))
        return s.span(L["qkv_b"], "grad", (3 * I,))

    # ---- geometry --------------------------------------------------------------------------------
    def geometry(self, B: int):
        np_, c = self.cfg.n_patches, self.cfg
        if self.mode == 0:
            lens = [np_] * B + [np_ + 1] * B
        elif self.mode == 1:
            lens = [np_ + 1] * B
        else:
            lens = [np_ + 1] * (2 * B)
        return lens

    def _buffers(self, B: int) -> _Buffers:
        bf = self.buf.get(B)
        if bf is not None:
            return bf
        bf = _Buffers(self.dev)
        lens = self.geometry(B)
        starts = [0]
        for n in lens:
            starts.append(starts[-1] + n)
        bf.T = starts[-1]
        bf.nseq = len(lens)
        bf.max_len = max(lens)
        bf.seq_start = torch.tensor(starts, dtype=torch.int32, device=self.dev)
        rows = starts[:-1]                     # token 0 of every sequence (x[:, 0], vit.py:162 / model.py:316)
        if self.mode == 1:
            rows = rows + rows                 # CLS feature feeds both value heads (model.py:300-302)
        bf.pool_rows = torch.tensor(rows, dtype=torch.int32, device=self.dev)
        bf.first_rows = torch.tensor(starts[:-1], dtype=torch.int32, device=self.dev)     # token 0 of every sequence, once
        nseq = len(lens)
        bf.pool_map = torch.tensor([i % nseq for i in range(2 * B)], dtype=torch.int32, device=self.dev)  # feature row -> sequence
        self.buf[B] = bf
        return bf

    # ---- forward ---------------------------------------------------------------------------------
    # dropout sites (vit.py:31,33,45,56,158): one well-mixed seed per (forward call, layer, site); the backward
    # regenerates the masks from the seeds saved in the per-B buffers
    SITE_EMB, SITE_ATTN_P, SITE_ATTN_OUT, SITE_ACT, SITE_FF_OUT = 0, 1, 2, 3, 4

    def _site(self, bf, li: int, kind: int):
        """(p, seed) of a dropout site for the forward call recorded in ``bf`` (p = 0 when dropout is off)."""
        if bf.drop_base is None:
            return 0.0, 0
        c = self.cfg
        p = (c.emb_dropout, c.attn_dropout, c.dropout, c.act_dropout, c.dropout)[kind]
        return (p, ops.site_seed(bf.drop_base, li * 8 + kind)) if p > 0 else (0.0, 0)

    def forward(self, img: torch.Tensor, B: int, sample_idx: Optional[torch.Tensor] = None,
                drop_base: Optional[int] = None) -> torch.Tensor:
        """img: [N,C,H,W] uint8 (raw frames, divided by 255 in-kernel) or float32 (already /255).
        Returns the pooled, final-LayerNorm'ed features F fp32 [2B, D] (rows [0,B) explorative/CLS,
        rows [B,2B) exploitative/CLS).  Activations stay in the per-B buffers for ``backward``.
        ``drop_base``: None = dropout off (eval mode / p = 0); an integer = this call's dropout stream."""
        c, s, p = self.cfg, self.store, self.pre
        bf = self._buffers(B)
        bf.drop_base = drop_base
        T, D, I, np_, PD = bf.T, c.dim, self.inner, c.n_patches, c.patch_dim
        rows = B * np_
        img_dt = ops._DT[img.dtype]
        bf.img, bf.sample_idx, bf.B = img, sample_idx, B
        pln = bf.get("pln", (rows, PD), torch.bfloat16)
        e0 = bf.get("e0", (rows, D), torch.float32)
        x0 = bf.get("x0", (T, D), torch.float32)
        fused_embed = (c.impl == "lucidrains" and D == 256 and PD % 16 == 0 and PD <= 192 and c.channels * c.patch * c.image <= 2080
                       and os.environ.get("EAVIT_FUSE_EMBED", "1") == "1")
        bf.pln_is_xhat = bool(fused_embed and self.fuse_embed_bwd)      # which backward of the patch Linear matches `pln`
        if fused_embed:
            # patchify + LayerNorm(PD) + Linear + LayerNorm(D) + token / position assembly: one kernel (csrc/embed_fused.cu)
            pm, pr = bf.get("pmean", (rows,), torch.float32), bf.get("prstd", (rows,), torch.float32)
            m3, r3 = bf.get("m3", (rows,), torch.float32), bf.get("r3", (rows,), torch.float32)
            tokA = s.w(p + ("exploration_token" if c.use_explorative else "cls_token"))
            call("eavit_embed_fused_fwd", img, img_dt, sample_idx, B, c.channels, c.image, c.patch, self.mode,
                 s.w(p + "to_patch_embedding.1.weight"), s.w(p + "to_patch_embedding.1.bias"), 1e-5,
                 s.b16(p + "to_patch_embedding.2.weight"), s.w(p + "to_patch_embedding.2.bias"),
                 s.w(p + "to_patch_embedding.3.weight"), s.w(p + "to_patch_embedding.3.bias"), 1e-5,
                 s.w(p + "pos_embedding"), tokA, pln, pm, pr, e0, m3, r3, x0, int(self.fuse_embed_bwd))
        elif c.impl == "lucidrains":
            pm, pr = bf.get("pmean", (rows,), torch.float32), bf.get("prstd", (rows,), torch.float32)
            call("eavit_patchify", img, img_dt, sample_idx, B, c.channels, c.image, c.patch, 0,
                 s.w(p + "to_patch_embedding.1.weight"), s.w(p + "to_patch_embedding.1.bias"), 1e-5, pln, pm, pr)
            linear_fwd(pln, s.b16(p + "to_patch_embedding.2.weight"), bias=s.w(p + "to_patch_embedding.2.bias"), out_f32=e0)
            e1 = bf.get("e1", (rows, D), torch.float32)
            m3, r3 = bf.get("m3", (rows,), torch.float32), bf.get("r3", (rows,), torch.float32)
            call("eavit_layernorm_fwd", e0, D, s.w(p + "to_patch_embedding.3.weight"), s.w(p + "to_patch_embedding.3.bias"),
                 e1, F32, D, m3, r3, rows, D, 1e-5)
            tokA = s.w(p + ("exploration_token" if c.use_explorative else "cls_token"))
            call("eavit_embed_assemble", e1, s.w(p + "pos_embedding"), tokA, None, self.mode, B, np_, D, x0)
        else:
            e = p + "embeddings."
            call("eavit_patchify", img, img_dt, sample_idx, B, c.channels, c.image, c.patch, 1, None, None, 0.0, pln, None, None)
            w16 = s.b16(e + "patch_embeddings.projection.weight").view(D, PD)
            linear_fwd(pln, w16, bias=s.w(e + "patch_embeddings.projection.bias"), out_f32=e0)
            call("eavit_embed_assemble", e0, s.w(e + "position_embeddings"), s.w(e + "exploration_token"),
                 s.w(e + "exploitation_token"), 2, B, np_, D, x0)
        pe, se = self._site(bf, 0, self.SITE_EMB)
        if pe > 0:
            call("eavit_dropout_apply", x0, D, None, 0, x0, D, T, D, pe, se)
        x = x0
        # LayerNorm fused into the GEMM that produces its input row (out-proj + residual -> LN2, MLP2 + residual -> the next
        # layer's LN1): possible when one 256-column tile holds the whole row
        fuse_ln = D == 256 and os.environ.get("EAVIT_FUSE_LN", "1") == "1"
        ln1_done = False
        for li, L in enumerate(self.L):
            xn1 = bf.get(f"xn1_{li}", (T, D), torch.bfloat16)
            m1, r1 = bf.get(f"m1_{li}", (T,), torch.float32), bf.get(f"r1_{li}", (T,), torch.float32)
            if not ln1_done:
                call("eavit_layernorm_fwd", x, D, s.w(L["ln1"][0]), s.w(L["ln1"][1]), xn1, BF16, D, m1, r1, T, D, c.ln_eps)
            ln1_done = False
            qkv = bf.get(f"qkv_{li}", (T, 3 * I), torch.bfloat16)
            linear_fwd(xn1, self._qkv(L, "w16"), bias=self._qkv(L, "b"), out_bf16=qkv)
            if self.prune_last and li == c.depth - 1:
                x = self._last_layer_pooled_fwd(bf, li, L, x, qkv)
                break
            o = bf.get(f"o_{li}", (T, I), torch.bfloat16)
            lse = bf.get(f"lse_{li}", (T, c.heads), torch.float32)
            pa, sa = self._site(bf, li, self.SITE_ATTN_P)
            ops.attention_fwd(qkv, bf.seq_start, bf.nseq, bf.max_len, c.heads, c.dim_head, float(c.dim_head) ** -0.5, o, lse,
                              drop_p=pa, drop_seed=sa)
            xmid = bf.get(f"xmid_{li}", (T, D), torch.float32)
            po, so = self._site(bf, li, self.SITE_ATTN_OUT)
            xn2 = bf.get(f"xn2_{li}", (T, D), torch.bfloat16)
            m2, r2 = bf.get(f"m2_{li}", (T,), torch.float32), bf.get(f"r2_{li}", (T,), torch.float32)
            if fuse_ln:
                linear_fwd(o, s.b16(L["o_w"]), bias=s.w(L["o_b"]), residual=x, out_f32=xmid, out_bf16=xn2, drop_p=po, drop_seed=so,
                           ln=(s.w(L["ln2"][0]), s.w(L["ln2"][1]), m2, r2, c.ln_eps))
            else:
                linear_fwd(o, s.b16(L["o_w"]), bias=s.w(L["o_b"]), residual=x, out_f32=xmid, drop_p=po, drop_seed=so)
                call("eavit_layernorm_fwd", xmid, D, s.w(L["ln2"][0]), s.w(L["ln2"][1]), xn2, BF16, D, m2, r2, T, D, c.ln_eps)
            hpre = bf.get(f"hpre_{li}", (T, c.mlp_dim), torch.bfloat16)
            hact = bf.get(f"hact_{li}", (T, c.mlp_dim), torch.bfloat16)
            ph, sh = self._site(bf, li, self.SITE_ACT)
            linear_fwd(xn2, s.b16(L["w1"]), bias=s.w(L["b1"]), act=ops.ACT_GELU_SAVE_GRAD, out_bf16=hact, out_pre=hpre, drop_p=ph, drop_seed=sh)
            xo = bf.get(f"x_{li + 1}", (T, D), torch.float32)
            pf, sf = self._site(bf, li, self.SITE_FF_OUT)
            if fuse_ln and li + 1 < c.depth:
                Ln = self.L[li + 1]
                linear_fwd(hact, s.b16(L["w2"]), bias=s.w(L["b2"]), residual=xmid, out_f32=xo, drop_p=pf, drop_seed=sf,
                           out_bf16=bf.get(f"xn1_{li + 1}", (T, D), torch.bfloat16),
                           ln=(s.w(Ln["ln1"][0]), s.w(Ln["ln1"][1]), bf.get(f"m1_{li + 1}", (T,), torch.float32),
                               bf.get(f"r1_{li + 1}", (T,), torch.float32), c.ln_eps))
                ln1_done = True
            else:
                linear_fwd(hact, s.b16(L["w2"]), bias=s.w(L["b2"]), residual=xmid, out_f32=xo, drop_p=pf, drop_seed=sf)
            x = xo
        nf = 2 * B
        pooled = bf.get("pooled", (nf, D), torch.float32)
        if self.prune_last:                                   # x = last layer's output for the pooled rows only [nseq, D]
            call("eavit_gather_rows", x, D, bf.pool_map, pooled, D, nf, D)
        else:
            call("eavit_gather_rows", x, D, bf.pool_rows, pooled, D, nf, D)
        feat = bf.get("feat", (nf, D), torch.float32)
        mf, rf = bf.get("mf", (nf,), torch.float32), bf.get("rf", (nf,), torch.float32)
        fn = (p + "transformer.norm.") if c.impl == "lucidrains" else (p + "layernorm.")
        call("eavit_layernorm_fwd", pooled, D, s.w(fn + "weight"), s.w(fn + "bias"), feat, F32, D, mf, rf, nf, D, c.ln_eps)
        return feat

    def _last_layer_pooled_fwd(self, bf, li, L, x, qkv):
        """Last layer after the QKV GEMM, for token 0 of every sequence only: [nseq, *] compact buffers."""
        c, s = self.cfg, self.store
        D, I, nc = c.dim, self.inner, bf.nseq
        o0 = bf.get("o0", (nc, I), torch.bfloat16)
        pa, sa = self._site(bf, li, self.SITE_ATTN_P)
        call("eavit_attention_row0_fwd", qkv, bf.seq_start, nc, bf.max_len, c.heads, c.dim_head, float(c.dim_head) ** -0.5, o0,
             pa, sa)
        xc = bf.get("xc", (nc, D), torch.float32)
        call("eavit_gather_rows", x, D, bf.first_rows, xc, D, nc, D)
        xmid = bf.get("xmid_c", (nc, D), torch.float32)
        po, so = self._site(bf, li, self.SITE_ATTN_OUT)
        linear_fwd(o0, s.b16(L["o_w"]), bias=s.w(L["o_b"]), residual=xc, out_f32=xmid, drop_p=po, drop_seed=so)
        xn2 = bf.get("xn2_c", (nc, D), torch.bfloat16)
        m2, r2 = bf.get("m2_c", (nc,), torch.float32), bf.get("r2_c", (nc,), torch.float32)
        call("eavit_layernorm_fwd", xmid, D, s.w(L["ln2"][0]), s.w(L["ln2"][1]), xn2, BF16, D, m2, r2, nc, D, c.ln_eps)
        hpre = bf.get("hpre_c", (nc, c.mlp_dim), torch.bfloat16)
        hact = bf.get("hact_c", (nc, c.mlp_dim), torch.bfloat16)
        ph, sh = self._site(bf, li, self.SITE_ACT)
        linear_fwd(xn2, s.b16(L["w1"]), bias=s.w(L["b1"]), act=ops.ACT_GELU_SAVE_GRAD, out_bf16=hact, out_pre=hpre, drop_p=ph, drop_seed=sh)
        xl = bf.get("xl_c", (nc, D), torch.float32)
        pf, sf = self._site(bf, li, self.SITE_FF_OUT)
        linear_fwd(hact, s.b16(L["w2"]), bias=s.w(L["b2"]), residual=xmid, out_f32=xl, drop_p=pf, drop_seed=sf)
        return xl

    def _last_layer_pooled_bwd(self, bf, li, L, top, dqkv, dxa):
        """Backward of ``_last_layer_pooled_fwd``.  top fp32 [nseq, D] = gradient of the layer's pooled outputs.  Writes the
        dense dqkv [T, 3I] and returns the pooled rows' residual gradient [nseq, D]."""
        c, s = self.cfg, self.store
        D, I, nc, T = c.dim, self.inner, bf.nseq, bf.T
        pf, sf = self._site(bf, li, self.SITE_FF_OUT)
        top_m = top
        if pf > 0:                                            # gradient of the MLP2 output: under its dropout mask
            top_m = bf.get("top_drop_c", (nc, D), torch.float32)
            call("eavit_dropout_apply", top, D, None, 0, top_m, D, nc, D, pf, sf)
        call("eavit_colsum", top_m, F32, D, s.g(L["b2"]), nc, D)
        top16 = bf.get("top16_c", (nc, D), torch.bfloat16)
        call("eavit_cast_f32_bf16", top_m, top16, nc * D)
        dh = bf.get("dh_c", (nc, c.mlp_dim), torch.bfloat16)
        ph, sh = self._site(bf, li, self.SITE_ACT)
        linear_bwd(top16, bf.t["hact_c"], s.b16(L["w2"]), dW=s.g(L["w2"]), db=None, dx_bf16=dh, act=ops.ACT_MUL_AUX,
                   aux=bf.t["hpre_c"], dx_colsum=s.g(L["b1"]), drop_p=ph, drop_seed=sh)
        dxn = bf.get("dxn_c", (nc, D), torch.bfloat16)
        linear_bwd(dh, bf.t["xn2_c"], s.b16(L["w1"]), dW=s.g(L["w1"]), db=None, dx_bf16=dxn)
        dxc = bf.get("dxc", (nc, D), torch.float32)
        dx16 = bf.get("dx16_c", (nc, D), torch.bfloat16)
        po, so = self._site(bf, li, self.SITE_ATTN_OUT)
        call("eavit_layernorm_bwd", dxn, BF16, D, bf.t["xmid_c"], D, bf.t["m2_c"], bf.t["r2_c"], s.w(L["ln2"][0]),
             top, D, dxc, D, dx16, D, s.g(L["ln2"][0]), s.g(L["ln2"][1]), s.g(L["o_b"]), po, so, nc, D)
        do0 = bf.get("do0", (nc, I), torch.bfloat16)
        linear_bwd(dx16, bf.t["o0"], s.b16(L["o_w"]), dW=s.g(L["o_w"]), db=None, dx_bf16=do0)
        pa, sa = self._site(bf, li, self.SITE_ATTN_P)
        call("eavit_attention_row0_bwd", bf.t[f"qkv_{li}"], do0, bf.seq_start, nc, bf.max_len, c.heads, c.dim_head,
             float(c.dim_head) ** -0.5, dqkv, pa, sa)
        return dxc            # [nseq, D] residual gradient of the pooled rows: added sparsely after the LayerNorm backward

    # ---- backward --------------------------------------------------------------------------------
    def layer_param_names(self, li: int) -> List[str]:
        """Every tensor of transformer layer ``li`` (its gradients are complete once ``backward`` has enqueued that layer)."""
        out = []
        for v in self.L[li].values():
            if v is None:
                continue
            out.extend(v if isinstance(v, (list, tuple)) else [v])
        return out

    def backward(self, dfeat: torch.Tensor, on_layer_done=None):
        """dfeat fp32 [2B, D]; accumulates every parameter gradient into ``store.grad``.
        ``on_layer_done(li)`` is called right after the last kernel that writes a gradient of layer ``li`` has been enqueued
        (reverse layer order) -- the data-parallel path starts that layer's slice of the gradient exchange there."""
        c, s, p = self.cfg, self.store, self.pre
        B = dfeat.shape[0] // 2
        bf = self.buf[B]
        T, D, I, np_, PD = bf.T, c.dim, self.inner, c.n_patches, c.patch_dim
        nf = 2 * B
        fn = (p + "transformer.norm.") if c.impl == "lucidrains" else (p + "layernorm.")
        dpool = bf.get("dpool", (nf, D), torch.float32)
        call("eavit_layernorm_bwd", dfeat, F32, D, bf.t["pooled"], D, bf.t["mf"], bf.t["rf"], s.w(fn + "weight"),
             None, D, dpool, D, None, D, s.g(fn + "weight"), s.g(fn + "bias"), None, 0.0, 0, nf, D)
        dxa = bf.get("dxa", (T, D), torch.float32)
        dxb = bf.get("dxb", (T, D), torch.float32)
        dx16 = bf.get("dx16", (T, D), torch.bfloat16)
        # The top gradient is sparse (pooled rows only).
        top, ntop = dpool, nf
        if self.mode == 1:
            top, ntop = bf.get("dpool_sum", (B, D), torch.float32), B
            call("eavit_add_f32", dpool[:B], dpool[B:], top, B * D)
        if not self.prune_last:
            # dense path: dxa = gradient of the residual stream; dx16 / db2 = gradient of the last MLP2 output = the same
            # rows under that layer's output-dropout mask (vit.py:33)
            call("eavit_zero", dxa, T * D * 4)
            call("eavit_zero", dx16, T * D * 2)
            pf, sf = self._site(bf, c.depth - 1, self.SITE_FF_OUT)
            top16 = top
            if pf > 0:
                top16 = bf.get("dpool_drop", (ntop, D), torch.float32)
                call("eavit_dropout_apply", top, D, bf.pool_rows, 0, top16, D, ntop, D, pf, sf)
            # bias gradient of the last layer's MLP2 = column sums of its (sparse) output gradient
            call("eavit_colsum", top16, F32, D, s.g(self.L[-1]["b2"]), ntop, D)
            if top16 is top:
                call("eavit_scatter_rows", top, D, bf.pool_rows, dxa, D, dx16, D, ntop, D)
            else:
                call("eavit_scatter_rows", top, D, bf.pool_rows, dxa, D, None, D, ntop, D)
                call("eavit_scatter_rows", top16, D, bf.pool_rows, None, D, dx16, D, ntop, D)
        dx, dx_other = dxa, dxb
        dh = bf.get("dh", (T, c.mlp_dim), torch.bfloat16)
        dxn = bf.get("dxn", (T, D), torch.bfloat16)     # LN-backward input (a GEMM output): bf16 halves its traffic
        do = bf.get("do", (T, I), torch.bfloat16)
        dqkv = bf.get("dqkv", (T, 3 * I), torch.bfloat16)
        # lucidrains, dim 256: the embedding backward is one pass that can also run layer 0's pre-attention LayerNorm backward on
        # the rows it reads (eavit_embed_assemble_ln_bwd, l1_* arguments) -- that LayerNorm's launch and the fp32 [T, D] gradient
        # between the two kernels disappear
        fuse_l1 = c.impl == "lucidrains" and D == 256 and self.fuse_embed_bwd and self.fuse_embed_ln1
        l1_args = None
        for li in reversed(range(c.depth)):
            L = self.L[li]
            x_in = bf.t["x0"] if li == 0 else bf.t[f"x_{li}"]
            if self.prune_last and li == c.depth - 1:
                # pooled rows only down to the attention, then the ordinary dense QKV / LN1 backward
                dxc = self._last_layer_pooled_bwd(bf, li, L, top, dqkv, dx)
                linear_bwd(dqkv, bf.t[f"xn1_{li}"], self._qkv(L, "w16"), dW=self._qkv(L, "gw"), db=self._qkv(L, "gb"), dx_bf16=dxn)
                db2_prev = s.g(self.L[li - 1]["b2"]) if li > 0 else None
                pf, sf = self._site(bf, li - 1, self.SITE_FF_OUT) if li > 0 else (0.0, 0)
                # the residual gradient entering this LayerNorm is zero except at the pooled rows: no dense `dres` stream
                # (206 MB zero fill + 206 MB read at cfg3); the pooled rows are added afterwards
                call("eavit_layernorm_bwd", dxn, BF16, D, x_in, D, bf.t[f"m1_{li}"], bf.t[f"r1_{li}"], s.w(L["ln1"][0]),
                     None, D, dx_other, D, dx16, D, s.g(L["ln1"][0]), s.g(L["ln1"][1]), db2_prev, pf, sf, T, D)
                call("eavit_scatter_add_rows", dxc, D, bf.first_rows, dx_other, D, dx16, D, db2_prev, bf.nseq, D, pf, sf)
                dx, dx_other = dx_other, dx
                if on_layer_done is not None:
                    on_layer_done(li)
                continue
            # MLP2: x_out = xmid + hact W2^T + b2
            # (db2 comes from the producer of dx: LN-bwd / top; db1 = colsum(dh) from this GEMM's epilogue)
            ph, sh = self._site(bf, li, self.SITE_ACT)
            linear_bwd(dx16, bf.t[f"hact_{li}"], s.b16(L["w2"]), dW=s.g(L["w2"]), db=None, dx_bf16=dh,
                       act=ops.ACT_MUL_AUX, aux=bf.t[f"hpre_{li}"], dx_colsum=s.g(L["b1"]), drop_p=ph, drop_seed=sh)
            # MLP1: hpre = xn2 W1^T + b1
            linear_bwd(dh, bf.t[f"xn2_{li}"], s.b16(L["w1"]), dW=s.g(L["w1"]), db=None, dx_bf16=dxn)
            po, so = self._site(bf, li, self.SITE_ATTN_OUT)     # dx16 / db_o: gradient of the out-proj output (masked)
            call("eavit_layernorm_bwd", dxn, BF16, D, bf.t[f"xmid_{li}"], D, bf.t[f"m2_{li}"], bf.t[f"r2_{li}"],
                 s.w(L["ln2"][0]), dx, D, dx_other, D, dx16, D, s.g(L["ln2"][0]), s.g(L["ln2"][1]), s.g(L["o_b"]), po, so, T, D)
            dx, dx_other = dx_other, dx
            # out-proj: xmid = x + o Wo^T + bo
            linear_bwd(dx16, bf.t[f"o_{li}"], s.b16(L["o_w"]), dW=s.g(L["o_w"]), db=None, dx_bf16=do)
            pa, sa = self._site(bf, li, self.SITE_ATTN_P)
            ops.attention_bwd(bf.t[f"qkv_{li}"], bf.t[f"o_{li}"], do, bf.t[f"lse_{li}"], bf.seq_start, bf.nseq,
                              bf.max_len, c.heads, c.dim_head, float(c.dim_head) ** -0.5, dqkv, drop_p=pa, drop_seed=sa)
            linear_bwd(dqkv, bf.t[f"xn1_{li}"], self._qkv(L, "w16"), dW=self._qkv(L, "gw"), db=self._qkv(L, "gb"), dx_bf16=dxn)
            db2_prev = s.g(self.L[li - 1]["b2"]) if li > 0 else None      # dx of this LN is the output gradient of layer li-1's MLP2
            pf, sf = self._site(bf, li - 1, self.SITE_FF_OUT) if li > 0 else (0.0, 0)   # dx16 / db2: layer li-1's MLP2 output
            if li == 0 and fuse_l1:
                # deferred into the embedding backward; `dx` stays the residual gradient at this LayerNorm's input
                l1_args = (dxn, x_in, bf.t["m1_0"], bf.t["r1_0"], s.w(L["ln1"][0]), s.g(L["ln1"][0]), s.g(L["ln1"][1]))
                continue
            # (layer 0 has no MLP2 below it: nobody reads the bf16 copy)
            call("eavit_layernorm_bwd", dxn, BF16, D, x_in, D, bf.t[f"m1_{li}"], bf.t[f"r1_{li}"], s.w(L["ln1"][0]),
                 dx, D, dx_other, D, dx16 if li > 0 else None, D, s.g(L["ln1"][0]), s.g(L["ln1"][1]), db2_prev, pf, sf, T, D)
            dx, dx_other = dx_other, dx
            if on_layer_done is not None:
                on_layer_done(li)
        # embedding (dropout after the positional add, vit.py:158)
        pe, se = self._site(bf, 0, self.SITE_EMB)
        fused_bwd = c.impl == "lucidrains" and D == 256 and self.fuse_embed_bwd
        if pe > 0 and not fused_bwd:                       # the fused backward reads dx under the mask itself
            call("eavit_dropout_apply", dx, D, None, 0, dx, D, T, D, pe, se)
        rows = B * np_
        img, sidx = bf.img, bf.sample_idx
        img_dt = ops._DT[img.dtype]
        if c.impl == "lucidrains":
            tok = p + ("exploration_token" if c.use_explorative else "cls_token")
            de16 = bf.get("de16", (rows, D), torch.bfloat16)
            fold = fused_bwd and getattr(bf, "pln_is_xhat", False)
            if fold:
                # `pln` holds xhat (no affine) and the frames need no gradient: G = de^T xhat and s = colsum(de) give the Linear's
                # weight / bias gradient AND LayerNorm(patch_dim)'s dgamma / dbeta (eavit_patch_ln_fold_bwd) -- no dX GEMM, no
                # second pass over the frames
                gs = bf.get("embed_G_s", (D * PD + D,), torch.float32)
                call("eavit_zero", gs, gs.numel() * 4)
                G, sv = gs[: D * PD].view(D, PD), gs[D * PD:]
            if fused_bwd:
                # token / position gradients, the sum over the two passes and the LayerNorm(dim) backward in one pass over dx
                call("eavit_embed_assemble_ln_bwd", dx, self.mode, B, np_, D, bf.t["e0"], bf.t["m3"], bf.t["r3"],
                     s.w(p + "to_patch_embedding.3.weight"), de16, s.g(p + "to_patch_embedding.3.weight"),
                     s.g(p + "to_patch_embedding.3.bias"), sv if fold else s.g(p + "to_patch_embedding.2.bias"),
                     s.g(p + "pos_embedding"), s.g(tok), None, pe, se, *(l1_args or (None,) * 7))
                if l1_args is not None and on_layer_done is not None:
                    on_layer_done(0)
            else:
                g = bf.get("g_embed", (rows, D), torch.float32)
                call("eavit_embed_assemble_bwd", dx, self.mode, B, np_, D, g, None, s.g(p + "pos_embedding"), s.g(tok), None)
                call("eavit_layernorm_bwd", g, F32, D, bf.t["e0"], D, bf.t["m3"], bf.t["r3"], s.w(p + "to_patch_embedding.3.weight"),
                     None, D, None, D, de16, D, s.g(p + "to_patch_embedding.3.weight"), s.g(p + "to_patch_embedding.3.bias"),
                     s.g(p + "to_patch_embedding.2.bias"), 0.0, 0, rows, D)
            if fold:
                ops.gemm(de16, bf.t["pln"], a_mn=True, b_mn=True, out_f32=G, atomic=True, split_k=_split_k(D, PD, rows))
                call("eavit_patch_ln_fold_bwd", G, sv, s.w(p + "to_patch_embedding.2.weight"), s.w(p + "to_patch_embedding.1.weight"),
                     s.w(p + "to_patch_embedding.1.bias"), s.g(p + "to_patch_embedding.2.weight"),
                     s.g(p + "to_patch_embedding.2.bias"), s.g(p + "to_patch_embedding.1.weight"),
                     s.g(p + "to_patch_embedding.1.bias"), D, PD)
            else:
                dpln = bf.get("dpln", (rows, PD), torch.float32)
                linear_bwd(de16, bf.t["pln"], s.b16(p + "to_patch_embedding.2.weight"), dW=s.g(p + "to_patch_embedding.2.weight"),
                           db=None, dx_f32=dpln)
                call("eavit_patchify_ln_bwd", img, img_dt, sidx, B, c.channels, c.image, c.patch, 0,
                     s.w(p + "to_patch_embedding.1.weight"), bf.t["pmean"], bf.t["prstd"], dpln,
                     s.g(p + "to_patch_embedding.1.weight"), s.g(p + "to_patch_embedding.1.bias"))
        else:
            e = p + "embeddings."
            g16 = bf.get("de16", (rows, D), torch.bfloat16)
            call("eavit_embed_assemble_bwd", dx, 2, B, np_, D, None, g16, s.g(e + "position_embeddings"),
                 s.g(e + "exploration_token"), s.g(e + "exploitation_token"))
            linear_bwd(g16, bf.t["pln"], None, dW=s.g(e + "patch_embeddings.projection.weight").view(D, PD),
                       db=s.g(e + "patch_embeddings.projection.bias"))


class Heads:
    """model.py:227-246 heads on the pooled features F [2B, D] (fp32 CUDA-core kernels: 0.02 % of the FLOPs)."""

    def __init__(self, cfg: HotPathConfig, store: ParamStore, n_actions: int, ext_uses_int_critic: bool):
        self.cfg, self.s, self.A = cfg, store, n_actions
        self.bug = ext_uses_int_critic      # model.py:321 (HG branch): value_ext = critic_int(...)
        self.buf: Dict[int, _Buffers] = {}
        self.coef = 0.5                     # attn_aggregation_op = 'mean' (model.py:284-286)

    def forward(self, feat: torch.Tensor):
        s, D, A = self.s, self.cfg.dim, self.A
        R = feat.shape[0]
        B = R // 2
        bf = self.buf.setdefault(B, _Buffers(feat.device))
        bf.feat = feat
        E = bf.get("E", (R, D), torch.float32)
        call("eavit_sgemm_small", feat, D, 0, s.w("model.extra_layer.0.weight"), D, 0, s.w("model.extra_layer.0.bias"), None,
             E, D, R, D, D, 1, 0)
        v = bf.get("v", (R,), torch.float32)
        wi, bi = s.w("model.critic_int.weight"), s.w("model.critic_int.bias")
        we, be = s.w("model.critic_ext.weight"), s.w("model.critic_ext.bias")
        if self.bug:
            call("eavit_heads_value_fwd", E, feat, wi, bi, wi, bi, R, R, D, v)
        else:
            call("eavit_heads_value_fwd", E, feat, wi, bi, we, be, B, R, D, v)
        comb = bf.get("comb", (B, D), torch.float32)
        call("eavit_combine_fwd", feat, comb, B, D, self.coef)
        a1 = bf.get("a1", (B, D), torch.float32)
        call("eavit_sgemm_small", comb, D, 0, s.w("model.actor.0.weight"), D, 0, s.w("model.actor.0.bias"), None, a1, D, B, D, D, 1, 0)
        pol = bf.get("policy", (B, A), torch.float32)
        call("eavit_sgemm_small", a1, D, 0, s.w("model.actor.2.weight"), D, 0, s.w("model.actor.2.bias"), None, pol, A, B, A, D, 0, 0)
        return pol, v[B:], v[:B]            # policy, value_ext, value_int

    def backward(self, dpol: torch.Tensor, dv: torch.Tensor) -> torch.Tensor:
        """dpol [B,A], dv [2B] = (dv_int | dv_ext) in feature-row order.  Returns dfeat [2B, D]."""
        s, D, A = self.s, self.cfg.dim, self.A
        B = dpol.shape[0]
        R = 2 * B
        bf = self.buf[B]
        feat, E, comb, a1 = bf.feat, bf.t["E"], bf.t["comb"], bf.t["a1"]
        dE, dF = bf.get("dE", (R, D), torch.float32), bf.get("dF", (R, D), torch.float32)
        wi, we = s.w("model.critic_int.weight"), s.w("model.critic_ext.weight")
        gi, ge = s.g("model.critic_int.weight"), s.g("model.critic_ext.weight")
        gbi, gbe = s.g("model.critic_int.bias"), s.g("model.critic_ext.bias")
        if self.bug:
            call("eavit_heads_value_bwd", E, feat, dv, wi, wi, R, R, D, dE, dF, gi, gbi, gi, gbi)
        else:
            call("eavit_heads_value_bwd", E, feat, dv, wi, we, B, R, D, dE, dF, gi, gbi, ge, gbe)
        # extra_layer: E = relu(F Wx^T + bx);  dE already masked
        call("eavit_sgemm_small", dE, D, 1, feat, D, 1, None, None, s.g("model.extra_layer.0.weight"), D, D, D, R, 0, 1)
        call("eavit_colsum", dE, F32, D, s.g("model.extra_layer.0.bias"), R, D)
        call("eavit_sgemm_small", dE, D, 0, s.w("model.extra_layer.0.weight"), D, 1, None, None, dF, D, R, D, D, 0, 1)
        # actor: policy = a1 Wa2^T + b ; a1 = relu(comb Wa0^T + b)
        call("eavit_sgemm_small", dpol, A, 1, a1, D, 1, None, None, s.g("model.actor.2.weight"), D, A, D, B, 0, 1)
        call("eavit_colsum", dpol, F32, A, s.g("model.actor.2.bias"), B, A)
        da1 = bf.get("da1", (B, D), torch.float32)
        call("eavit_sgemm_small", dpol, A, 0, s.w("model.actor.2.weight"), D, 1, None, a1, da1, D, B, D, A, 0, 0)
        call("eavit_sgemm_small", da1, D, 1, comb, D, 1, None, None, s.g("model.actor.0.weight"), D, D, D, B, 0, 1)
        call("eavit_colsum", da1, F32, D, s.g("model.actor.0.bias"), B, D)
        dcomb = bf.get("dcomb", (B, D), torch.float32)
        call("eavit_sgemm_small", da1, D, 0, s.w("model.actor.0.weight"), D, 1, None, None, dcomb, D, B, D, D, 0, 0)
        call("eavit_combine_bwd", dcomb, dF, B, D, self.coef)
        return dF


class ConvTower:
    """conv8s4 - act - conv4s2 - act - conv3s1 - act - Flatten - Linear [- ReLU - Linear ...] as im2col + tcgen05 GEMMs.

    Two users: the RND towers (model.py:366-416: 1 input channel, LeakyReLU, Linear(3136,512) [+ 2 x Linear(512,512)], no final
    activation) and the CNN actor-critic backbone of BASELINE configs[1] (model.py:110-135, commented out upstream: the 4-frame
    NCHW stack, ReLU, Linear(3136,256)-ReLU-Linear(256,448)-ReLU).

    Forward operands are bf16x3 splits ([hi|hi|lo] x [hi|lo|hi], one GEMM with K' = 3K): the (Leaky)ReLU masks of
    these plain ReLU networks flip wherever a pre-activation's rounding error exceeds its magnitude, and with single
    bf16 operands (2^-9) those flips alone put ~3 % error on the gradients.  The towers are < 1 % of the
    update's FLOPs, so near-fp32 pre-activations cost nothing measurable.  Backward GEMMs use plain bf16 (hi parts).
    """

    def __init__(self, store: ParamStore, prefix: str, cin: int, slope: float, fcs: Sequence[int], fc_out: Sequence[int],
                 final_relu: bool, nchw_input: bool, image: int = 84):
        self.s, self.pre, self.image = store, prefix, image
        self.cin, self.slope, self.fcs, self.fc_out = cin, float(slope), tuple(fcs), tuple(fc_out)
        self.final_relu, self.nchw = final_relu, nchw_input
        self.conv_act = ops.ACT_LRELU if slope > 0 else ops.ACT_RELU
        self.CONVS = ((8, 4, cin, 32), (4, 2, 32, 64), (3, 1, 64, 64))   # (kernel, stride, Cin, Cout)
        self.out = self.fc_out[-1]
        self.buf: Dict[int, _Buffers] = {}
        h = image
        self.sizes = []
        for k, st, ci, co in self.CONVS:
            oh = (h - k) // st + 1
            self.sizes.append((h, oh))
            h = oh
        self.flat_dim = h * h * 64
        self.w3: Dict[str, torch.Tensor] = {}
        for idx in (0, 2, 4) + self.fcs:
            w = store.w(self.pre + f"{idx}.weight")
            n, k = w.shape[0], w[0].numel()
            self.w3[f"{idx}"] = torch.empty(n, 3 * k, dtype=torch.bfloat16, device=store.device)
        self.refresh_weights()

    def refresh_weights(self):
        """fp32 master weights -> [hi | lo | hi] bf16 GEMM operands (after load / every optimiser step)."""
        for idx, w3 in self.w3.items():
            w = self.s.w(self.pre + f"{idx}.weight")
            n, k = w.shape[0], w[0].numel()
            call("eavit_split3_rows", w, k, n, k, w3, 1)

    def forward(self, obs: torch.Tensor, B: int, sample_idx: Optional[torch.Tensor] = None,
                col0: Optional[torch.Tensor] = None) -> torch.Tensor:
        """obs: [N,1,H,W] fp32 (RND: normalised, clipped) or, with ``nchw_input``, the frame stack [N,C,H,W] uint8 / fp32;
        returns features fp32 [B, out].
        ``col0``: the first convolution's patch matrix of the SAME observations, already built by the other tower
        (``self.buf[B].t["col0"]`` of that tower) -- predictor and target read identical inputs (model.py:457-461)."""
        s, p = self.s, self.pre
        bf = self.buf.setdefault(B, _Buffers(obs.device))
        assert obs.dtype == torch.float32 or (self.nchw and obs.dtype == torch.uint8)
        x, sidx = obs, sample_idx
        for ci, ((k, st, cin, cout), (h, oh)) in enumerate(zip(self.CONVS, self.sizes)):
            K = cin * k * k
            if ci == 0 and col0 is not None:
                col = bf.t["col0"] = col0
            else:
                col = bf.get(f"col{ci}", (B * oh * oh, 3 * K), torch.bfloat16)
                if ci == 0 and self.nchw:
                    call("eavit_im2col_nchw", x, ops._DT[x.dtype], sidx, B, h, h, cin, k, k, st, col, 1)
                else:
                    call("eavit_im2col", x, F32, sidx, B, h, h, cin, k, k, st, col, 1)
            a32 = bf.get(f"act32_{ci}", (B * oh * oh, cout), torch.float32)
            a16 = bf.get(f"act{ci}", (B * oh * oh, cout), torch.bfloat16)
            ops.gemm(col, self.w3[f"{2 * ci}"], bias=s.w(p + f"{2 * ci}.bias"), act=self.conv_act, out_f32=a32, out_bf16=a16)
            x, sidx = a32, None
        hw = self.sizes[-1][1] ** 2
        flat32 = bf.get("flat32", (B, self.flat_dim), torch.float32)
        call("eavit_nhwc_to_flat_f32", x, B, hw, 64, flat32)
        h3 = bf.get("flat", (B, 3 * self.flat_dim), torch.bfloat16)
        call("eavit_split3_rows", flat32, self.flat_dim, B, self.flat_dim, h3, 0)
        out = None
        for j, k in enumerate(self.fcs):
            # M = B rows give only B/128 x 2 output tiles: split K over the idle SMs (atomic fp32 accumulation into a zeroed
            # buffer), then one small kernel adds the bias, applies the ReLU and emits the next layer's bf16x3 rows
            last = j == len(self.fcs) - 1
            bias = s.w(p + f"{k}.bias")
            w3 = self.w3[f"{k}"]
            n_out = self.fc_out[j]
            f32 = bf.get("out" if last else f"fc32_{j}", (B, n_out), torch.float32)
            call("eavit_zero", f32, f32.numel() * 4)
            ops.gemm(h3, w3, out_f32=f32, atomic=True, split_k=_split_k(B, n_out, h3.shape[1]))
            if last:
                call("eavit_bias_act_split3", f32, n_out, bias, ops.ACT_RELU if self.final_relu else ops.ACT_NONE, None, None, B, n_out)
                out = f32
            else:
                f16 = bf.get(f"fc16_{j}", (B, n_out), torch.bfloat16)
                h3 = bf.get(f"fc{j}", (B, 3 * n_out), torch.bfloat16)
                call("eavit_bias_act_split3", f32, n_out, bias, ops.ACT_RELU, f16, h3, B, n_out)
        return out

    def backward(self, dout: torch.Tensor):
        """dout [B, out]: bf16, or fp32 when the tower ends with a ReLU (masked here by the stored output);
        accumulates the tower's parameter gradients into ``store.grad``."""
        s, p = self.s, self.pre
        B = dout.shape[0]
        bf = self.buf[B]
        if self.final_relu:
            d = bf.get("dout16", (B, self.out), torch.bfloat16)
            call("eavit_act_bwd_bf16", dout, bf.t["out"], 0.0, d, B * self.out)
        else:
            d = dout
        nfc = len(self.fcs)
        for j in reversed(range(nfc)):
            name = p + f"{self.fcs[j]}."
            x16 = bf.t["flat"][:, : self.flat_dim] if j == 0 else bf.t[f"fc{j - 1}"][:, : self.fc_out[j - 1]]   # hi parts
            if j == 0:
                dflat = bf.get("dflat", (B, self.flat_dim), torch.bfloat16)
                linear_bwd(d, x16, s.b16(name + "weight"), dW=s.g(name + "weight"), db=s.g(name + "bias"), dx_bf16=dflat)
                d = dflat
            else:
                dprev = bf.get(f"dfc{j - 1}", (B, self.fc_out[j - 1]), torch.bfloat16)
                linear_bwd(d, x16, s.b16(name + "weight"), dW=s.g(name + "weight"), db=s.g(name + "bias"), dx_bf16=dprev,
                           act=ops.ACT_RELU_BWD, aux=bf.t[f"fc16_{j - 1}"])
                d = dprev
        hw = self.sizes[-1][1] ** 2
        dact = bf.get("dact2", (B * hw, 64), torch.bfloat16)
        call("eavit_flat_to_nhwc_act", d, bf.t["act2"], B, hw, 64, dact, self.slope)
        for ci in (2, 1, 0):
            k, st, cin, cout = self.CONVS[ci]
            h, oh = self.sizes[ci]
            K = cin * k * k
            name = p + f"{2 * ci}."
            col_hi = bf.t[f"col{ci}"][:, :K]
            if ci == 0:
                linear_bwd(dact, col_hi, None, dW=s.g(name + "weight").view(cout, K), db=s.g(name + "bias"))
                break
            dcol = bf.get(f"dcol{ci}", (B * oh * oh, K), torch.bfloat16)
            linear_bwd(dact, col_hi, s.b16(name + "weight").view(cout, K), dW=s.g(name + "weight").view(cout, K),
                       db=s.g(name + "bias"), dx_bf16=dcol)
            dprev = bf.get(f"dact{ci - 1}", (B * h * h, cin), torch.bfloat16)
            call("eavit_col2im_act", dcol, bf.t[f"act{ci - 1}"], B, h, h, cin, k, k, st, dprev, self.slope)
            dact = dprev


class RNDNet(ConvTower):
    """model.py:366-416 conv tower + FC stack; ``net`` = 'predictor' (trainable) or 'target' (frozen)."""

    def __init__(self, store: ParamStore, net: str, image: int = 84, out: int = 512):
        self.net = net
        fcs = (7, 9, 11) if net == "predictor" else (7,)
        super().__init__(store, f"rnd.{net}.", cin=1, slope=0.01, fcs=fcs, fc_out=(out,) * len(fcs), final_relu=False,
                         nchw_input=False, image=image)


class CnnEncoder:
    """The original RND CNN backbone (model.py:110-135, commented out upstream; BASELINE configs[1]) behind the encoder
    interface the heads expect: ``forward`` returns F fp32 [2B, D] with rows [0,B) == rows [B,2B) == the one feature vector of
    each sample (policy = actor(x), value = critic(extra_layer(x) + x) for both critics -- the CLS-style head wiring)."""

    def __init__(self, cfg: HotPathConfig, store: ParamStore, prefix: str = "model.feature."):
        self.cfg, self.store = cfg, store
        self.tower = ConvTower(store, prefix, cin=cfg.channels, slope=0.0, fcs=(7, 9), fc_out=(256, cfg.dim), final_relu=True,
                               nchw_input=True, image=cfg.image)
        self.buf: Dict[int, _Buffers] = {}

    def refresh_weights(self):
        self.tower.refresh_weights()

    def forward(self, img: torch.Tensor, B: int, sample_idx: Optional[torch.Tensor] = None, drop_base=None) -> torch.Tensor:
        x = self.tower.forward(img, B, sample_idx)
        bf = self.buf.setdefault(B, _Buffers(img.device))
        feat = bf.get("feat", (2 * B, self.cfg.dim), torch.float32)
        feat[:B].copy_(x)
        feat[B:].copy_(x)
        return feat

    def backward(self, dfeat: torch.Tensor):
        B = dfeat.shape[0] // 2
        bf = self.buf[B]
        d = bf.get("dx", (B, self.cfg.dim), torch.float32)
        call("eavit_add_f32", dfeat[:B], dfeat[B:], d, B * self.cfg.dim)
        self.tower.backward(d)
