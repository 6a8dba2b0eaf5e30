"""Drop-in for the reference's ``model.py``: ``CnnActorCriticNetwork``, ``RNDModel``, ``ViT_IMPLEMENTATION``.

Constructor signatures, attribute names, parameter names / shapes / init order follow model.py:85-263 and
:357-455, so ``state_dict`` round-trips with the reference.  ``forward`` runs on the sm_100a kernels through
``Runtime`` (flat parameter store + ``engine`` orchestration) and is differentiable: the custom autograd node
accumulates parameter gradients straight into ``p.grad`` (views of the flat gradient buffer).
"""
from __future__ import annotations

import math
import os
from collections import OrderedDict
from enum import Enum
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn
from torch.nn import init

from . import ops
from .config import HotPathConfig, default_config
from .engine import CnnEncoder, Heads, ParamStore, RNDNet, ViTEncoder
from .ops import call
from .utils import Env_action_space_type
from .vit import ViT, ViT_Attn
from .vit_hg import ViT_ExplorativeAttn, ViTConfigLike


class ViT_IMPLEMENTATION(Enum):   # model.py:16-18
    LUCIDRAINS_ViT = 0
    HG_ViT = 1
    # not in the reference's enum: selects the original RND CNN backbone that model.py:110-178 keeps as commented-out code
    # (BASELINE configs[1]); ``ViT_implementation_type = 2`` in the .conf
    ORIGINAL_CNN = 2


class Flatten(nn.Module):         # model.py:80-82 (parameter-free; kept for state_dict index alignment)
    def forward(self, input):
        return input.view(input.size(0), -1)


# --------------------------------------------------------------------------------------------------
# runtime: flat stores + engines for a module tree
# --------------------------------------------------------------------------------------------------
def _hg_adjacent_order(names: List[str]) -> List[str]:
    """Place HF query|key|value weights (and biases) next to each other so one GEMM covers all three."""
    out, used = [], set()
    for n in names:
        if n in used:
            continue
        if n.endswith("attention.attention.query.weight"):
            base = n[: -len("query.weight")]
            grp = [base + f"{k}.weight" for k in ("query", "key", "value")] + [base + f"{k}.bias" for k in ("query", "key", "value")]
            for g in grp:
                if g in names:
                    out.append(g)
                    used.add(g)
        else:
            out.append(n)
            used.add(n)
    return out


class Runtime:
    """Owns the flat parameter stores of ``root`` and the engines that consume them."""

    def __init__(self, root: nn.Module, prefix: str, n_actions: Optional[int] = None, ext_uses_int_critic: bool = False):
        named = [(prefix + n, p) for n, p in root.named_parameters()]
        assert named, "module has no parameters"
        dev = named[0][1].device
        if dev.type != "cuda":
            raise RuntimeError("eavit_b200 computes on CUDA (sm_100a) only -- move the module to a CUDA device first; "
                               "there is no CPU fallback")
        self.device = dev
        self.params = OrderedDict(named)
        train_names = [n for n, _ in named if n.startswith("model.") or n.startswith("rnd.predictor.")]
        frozen_names = [n for n, _ in named if n.startswith("rnd.target.")]
        train_names = _hg_adjacent_order(train_names)
        self.store = self._flatten(train_names, True)
        self.frozen = self._flatten(frozen_names, False) if frozen_names else None
        self.encoder = self.heads = self.rnd_pred = self.rnd_tgt = None
        feat = None
        for m in root.modules():
            if isinstance(m, (ViT, ViT_ExplorativeAttn)):
                feat = m
                break
        if feat is not None:
            hp = feat.hp
            self.cfg = HotPathConfig(impl="lucidrains" if isinstance(feat, ViT) else "hg", image=hp["image"], channels=hp["channels"],
                                     patch=hp["patch"], dim=hp["dim"], depth=hp["depth"], heads=hp["heads"], dim_head=hp["dim_head"],
                                     mlp_dim=hp["mlp_dim"], use_explorative=feat.use_explorativeAttn,
                                     ln_eps=(1e-5 if isinstance(feat, ViT) else feat.config.layer_norm_eps),
                                     dropout=hp["dropout"], emb_dropout=hp["emb_dropout"],
                                     attn_dropout=hp.get("attn_dropout", -1.0), act_dropout=hp.get("act_dropout", -1.0))
            self.encoder = ViTEncoder(self.cfg, self.store, "model.feature.")
            self._feat = feat
            # dropout stream: one counter per Runtime, offset by torch's seed and the rank (ranks draw independent masks)
            self._drop_seed0 = (torch.initial_seed() * 1000003 + int(os.environ.get("RANK", "0")) * 7919) & ((1 << 48) - 1)
            self._drop_calls = 0
            if "model.actor.0.weight" in self.params:
                A = self.params["model.actor.2.weight"].shape[0] if n_actions is None else n_actions
                self.heads = Heads(self.cfg, self.store, A, ext_uses_int_critic)
        elif "model.feature.0.weight" in self.params and "model.feature.9.weight" in self.params:
            # the original RND CNN backbone (model.py:110-135): no ViT, no dropout
            w0, w9 = self.params["model.feature.0.weight"], self.params["model.feature.9.weight"]
            self.cfg = HotPathConfig(impl="cnn", image=84, channels=w0.shape[1], patch=84, dim=w9.shape[0], depth=0, heads=0,
                                     dim_head=0, mlp_dim=0, use_explorative=False, ln_eps=0.0, dropout=0.0, emb_dropout=0.0)
            self.encoder = CnnEncoder(self.cfg, self.store, "model.feature.")
            self._feat = root
            self._drop_seed0 = self._drop_calls = 0
            A = self.params["model.actor.2.weight"].shape[0] if n_actions is None else n_actions
            self.heads = Heads(self.cfg, self.store, A, False)
        if "rnd.predictor.0.weight" in self.params:
            self.rnd_pred = RNDNet(self.store, "predictor")
            self.rnd_tgt = RNDNet(self.frozen, "target")
        self._first = named[0][1]

    def _flatten(self, names: List[str], trainable: bool) -> ParamStore:
        shapes = OrderedDict((n, tuple(self.params[n].shape)) for n in names)
        st = ParamStore(shapes, self.device, trainable)
        with torch.no_grad():
            for n in names:
                p = self.params[n]
                st.w(n).copy_(p.data)
                p.data = st.w(n)
                if trainable:
                    p.grad = st.g(n)
        st.sync_shadow()
        return st

    def valid(self) -> bool:
        n0 = next(iter(self.params))
        st = self.store if n0 in self.store.shapes else self.frozen
        return self._first.data_ptr() == st.w(n0).data_ptr()

    def _param_versions(self) -> int:
        """Sum of the autograd version counters of every parameter: load_state_dict, torch optimisers and any in-place op
        on a Parameter bump them; the fused Adam step (which maintains the shadows itself) does not."""
        return sum(p._version for p in self.params.values())

    def sync_if_changed(self):
        """``sync()`` only when a parameter was written from outside since the last sync -- the rollout calls this once per
        env step (12 launches saved per call).  Writes through ``p.data`` bypass the version counters: call ``sync()``."""
        v = self._param_versions()
        if v != getattr(self, "_synced_version", None):
            self.sync()

    def sync(self):
        """Refresh the bf16 shadows after the fp32 masters were written from outside (load_state_dict, ...)."""
        self._synced_version = self._param_versions()
        self.store.sync_shadow()
        if self.frozen is not None:
            self.frozen.sync_shadow()
        if self.rnd_pred is not None:
            self.rnd_pred.refresh_weights()
            self.rnd_tgt.refresh_weights()
        if isinstance(self.encoder, CnnEncoder):
            self.encoder.refresh_weights()
        # gradients may have been detached by optimizer.zero_grad(set_to_none=True)
        for n, p in self.params.items():
            if n in self.store.shapes and (p.grad is None or p.grad.data_ptr() != self.store.g(n).data_ptr()):
                p.grad = self.store.g(n)

    def refresh_after_step(self):
        """bf16x3 operand copies of the conv-tower weights follow the fp32 masters after every optimiser step (the bf16
        shadow of the flat store is written by the Adam kernel itself)."""
        if self.rnd_pred is not None:
            self.rnd_pred.refresh_weights()
        if isinstance(self.encoder, CnnEncoder):
            self.encoder.refresh_weights()

    # ---- actor-critic ----------------------------------------------------------------------------
    def dropout_active(self) -> bool:
        c = self.cfg
        return bool(self._feat.training and max(c.dropout, c.emb_dropout, c.attn_dropout, c.act_dropout) > 0
                    and os.environ.get("EAVIT_DROPOUT_AS_IDENTITY", "0") != "1")

    def next_drop_base(self):
        """Dropout stream id of the next forward call: None when dropout is off (eval mode or every p = 0).  Like the
        reference, dropout follows the module's train/eval flag -- the rollout runs in train mode too."""
        c = self.cfg
        if not self._feat.training or max(c.dropout, c.emb_dropout, c.attn_dropout, c.act_dropout) <= 0:
            return None
        if os.environ.get("EAVIT_DROPOUT_AS_IDENTITY", "0") == "1":
            return None
        self._drop_calls += 1
        return self._drop_seed0 + self._drop_calls

    def _bump_gen(self, kind: str, B: int) -> int:
        """Activations live in per-batch-size scratch buffers, not in the autograd graph: every forward of (kind, B)
        gets a generation number so that a backward whose buffers were overwritten since can refuse to run."""
        g = self._gens = getattr(self, "_gens", {})
        g[(kind, B)] = g.get((kind, B), 0) + 1
        return g[(kind, B)]

    def check_gen(self, kind: str, B: int, gen: int, epoch: int):
        if self._gens.get((kind, B)) != gen:
            raise RuntimeError(f"eavit_b200: backward of a {kind} forward (batch {B}) whose activations were overwritten by a later "
                               "forward of the same batch size -- run backward before the next forward (or use RNDAgent.train_step)")
        if epoch != getattr(self, "_epoch_gen", 0):
            raise RuntimeError("eavit_b200: a graph-replayed get_action() advanced the dropout epoch between this forward and its "
                               "backward; the regenerated masks would differ")

    def ac_forward(self, state: torch.Tensor, B: int, sample_idx=None):
        self._bump_gen("ac", B)
        feat = self.encoder.forward(state, B, sample_idx, drop_base=self.next_drop_base())
        return self.heads.forward(feat)          # policy [B,A], value_ext [B], value_int [B]  (views of scratch)

    def ac_backward(self, dpol: torch.Tensor, dv: torch.Tensor, backbone: bool = True, on_layer_done=None):
        """dv fp32 [2B] = (d value_int | d value_ext).  ``backbone=False``: the shared feature extractor is frozen
        (train.py:261-263) -- only the heads are differentiated.  ``on_layer_done``: see ViTEncoder.backward."""
        dfeat = self.heads.backward(dpol, dv)
        if backbone:
            if on_layer_done is not None and isinstance(self.encoder, ViTEncoder):
                self.encoder.backward(dfeat, on_layer_done)
            else:
                self.encoder.backward(dfeat)

    def frozen_names(self) -> List[str]:
        """Tensors of the trainable store whose Parameter has requires_grad = False (e.g. freeze_shared_backbone)."""
        return [n for n in self.store.shapes if not self.params[n].requires_grad]


def _runtime_for(module: nn.Module, prefix: str, **kw) -> Runtime:
    rt = getattr(module, "_rt", None)
    if rt is None or not rt.valid():
        rt = Runtime(module, prefix, **kw)
        module._rt = rt
    return rt


def _as_device_image(x, device) -> torch.Tensor:
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(x)
    if x.dtype not in (torch.uint8, torch.float32):
        x = x.float()
    return x.to(device).contiguous()


def standalone_vit_features(vit_module: nn.Module, img) -> torch.Tensor:
    """Pooled features [2B, D] of a bare ViT module (inference only)."""
    rt = _runtime_for(vit_module, "model.feature.")
    rt.sync()
    x = _as_device_image(img, rt.device)
    with torch.no_grad():
        return rt.encoder.forward(x, x.shape[0], drop_base=rt.next_drop_base()).clone()


class _ActorCriticFn(torch.autograd.Function):
    """(policy, value_ext, value_int) = f(state); backward accumulates into the flat gradient buffer."""

    @staticmethod
    def forward(ctx, anchor: torch.Tensor, rt: Runtime, state: torch.Tensor):
        B = state.shape[0]
        pol, ve, vi = rt.ac_forward(state, B)
        ctx.rt, ctx.B = rt, B
        ctx.gen, ctx.epoch = rt._gens[("ac", B)], getattr(rt, "_epoch_gen", 0)
        return pol.clone(), ve.clone().unsqueeze(1), vi.clone().unsqueeze(1)

    @staticmethod
    def backward(ctx, dpol, dve, dvi):
        rt, B = ctx.rt, ctx.B
        rt.check_gen("ac", B, ctx.gen, ctx.epoch)
        dev = rt.device
        dpol = torch.zeros(B, rt.heads.A, device=dev) if dpol is None else dpol.contiguous().float()
        dv = torch.zeros(2 * B, device=dev)
        if dvi is not None:
            dv[:B] = dvi.reshape(-1)
        if dve is not None:
            dv[B:] = dve.reshape(-1)
        rt.ac_backward(dpol, dv)
        return None, None, None


class CnnActorCriticNetwork(nn.Module):
    """model.py:85-354.  ViT backbone (lucidrains or HF style, from ``config.default_config``) + PPO heads."""

    extracted_feature_embedding_dim = int(default_config["extracted_feature_embedding_dim"])   # model.py:89

    def __init__(self, input_size, output_size, env_action_space_type, use_noisy_net=False,
                 ViT_implementation_type: ViT_IMPLEMENTATION = ViT_IMPLEMENTATION.LUCIDRAINS_ViT):
        super().__init__()
        assert isinstance(ViT_implementation_type, ViT_IMPLEMENTATION), "ViT_implementation_type must be of type enum ViT_IMPLEMENTATION"
        assert env_action_space_type == Env_action_space_type.DISCRETE, "hot path covers the DISCRETE branch (all ViT configs are Atari)"
        assert not use_noisy_net, "UseNoisyNet is False in every reference config (out of scope, SURVEY 2a row 3)"
        self.env_action_space_type = env_action_space_type
        self.ViT_implementation_type = ViT_implementation_type
        c = default_config
        if ViT_implementation_type == ViT_IMPLEMENTATION.ORIGINAL_CNN:         # model.py:110-135 (commented out upstream)
            ViT_dim = int(c["extracted_feature_embedding_dim"])
            self.feature = nn.Sequential(
                nn.Conv2d(in_channels=int(c["StateStackSize"]), out_channels=32, kernel_size=8, stride=4), nn.ReLU(),
                nn.Conv2d(in_channels=32, out_channels=64, kernel_size=4, stride=2), nn.ReLU(),
                nn.Conv2d(in_channels=64, out_channels=64, kernel_size=3, stride=1), nn.ReLU(),
                Flatten(), nn.Linear(7 * 7 * 64, 256), nn.ReLU(), nn.Linear(256, ViT_dim), nn.ReLU())
        elif ViT_implementation_type == ViT_IMPLEMENTATION.LUCIDRAINS_ViT:     # model.py:183-196
            ViT_dim = int(c["ViTlucidrains_dim"])
            self.feature = ViT(image_size=int(c["PreProcHeight"]), patch_size=int(c["ViTlucidrains_patch_size"]),
                               num_classes=int(c["ViTlucidrains_num_classes"]), dim=ViT_dim, depth=int(c["ViTlucidrains_depth"]),
                               heads=int(c["ViTlucidrains_heads"]), mlp_dim=int(c["ViTlucidrains_mlp_dim"]),
                               dropout=float(c["ViTlucidrains_dropout"]), emb_dropout=float(c["ViTlucidrains_emb_dropout"]),
                               channels=int(c["StateStackSize"]), dim_head=int(c["ViTlucidrains_dim_head"]),
                               use_explorativeAttn=c.getboolean("ViTlucidrains_use_explorativeAttn"))
        else:                                                                   # model.py:200-220
            ViT_dim = int(c["ViTHG_hidden_size"])
            cfg = ViTConfigLike(hidden_size=ViT_dim, num_hidden_layers=int(c["ViTHG_num_hidden_layers"]),
                                num_attention_heads=int(c["ViTHG_num_attention_heads"]),
                                intermediate_size=int(c["ViTHG_intermediate_size"]), hidden_act="gelu",
                                hidden_dropout_prob=float(c["ViTHG_hidden_dropout_prob"]),
                                attention_probs_dropout_prob=float(c["ViTHG_attention_probs_dropout_prob"]),
                                initializer_range=float(c["ViTHG_initializer_range"]), layer_norm_eps=float(c["ViTHG_layer_norm_eps"]),
                                image_size=int(c["ViTHG_PreProcHeight"]), patch_size=int(c["ViTHG_patch_size"]),
                                num_channels=int(c["ViTHG_StateStackSize"]), qkv_bias=c.getboolean("ViTHG_qkv_bias"),
                                encoder_stride=int(c["ViTHG_encoder_stride"]))
            self.feature = ViT_ExplorativeAttn(cfg, add_pooling_layer=True, use_mask_token=False,
                                               use_explorativeAttn=c.getboolean("ViTHG_use_explorativeAttn"))
            assert ViT_dim == int(c["extracted_feature_embedding_dim"]), \
                "In the provided config file 'VitHG_hidden_size' should equal 'extracted_feature_embedding_dim'."   # model.py:220
        self.actor = nn.Sequential(nn.Linear(ViT_dim, ViT_dim), nn.ReLU(), nn.Linear(ViT_dim, output_size))   # model.py:227-231
        self.extra_layer = nn.Sequential(nn.Linear(ViT_dim, ViT_dim), nn.ReLU())                                # model.py:240-243
        self.critic_ext = nn.Linear(ViT_dim, 1)
        self.critic_int = nn.Linear(ViT_dim, 1)
        if ViT_implementation_type == ViT_IMPLEMENTATION.ORIGINAL_CNN:         # model.py:147-155
            for p in self.modules():
                if isinstance(p, (nn.Conv2d, nn.Linear)):
                    init.orthogonal_(p.weight, np.sqrt(2))
                    p.bias.data.zero_()
        init.orthogonal_(self.critic_ext.weight, 0.01)                                                          # model.py:249-263
        self.critic_ext.bias.data.zero_()
        init.orthogonal_(self.critic_int.weight, 0.01)
        self.critic_int.bias.data.zero_()
        for i in range(len(self.actor)):
            if type(self.actor[i]) == nn.Linear:
                init.orthogonal_(self.actor[i].weight, 0.01)
                self.actor[i].bias.data.zero_()
        for i in range(len(self.extra_layer)):
            if type(self.extra_layer[i]) == nn.Linear:
                init.orthogonal_(self.extra_layer[i].weight, 0.1)
                self.extra_layer[i].bias.data.zero_()
        self.output_size = output_size
        self._rt = None
        self._agent_rt = None     # set by RNDAgent so model / rnd / optimiser share one flat store

    def runtime(self) -> Runtime:
        if self._agent_rt is not None:
            return self._agent_rt()
        return _runtime_for(self, "model.", n_actions=self.output_size,
                            ext_uses_int_critic=self.ViT_implementation_type == ViT_IMPLEMENTATION.HG_ViT)

    def forward(self, state, attn_aggregation_op="mean"):
        """model.py:266-352 -> (policy [B,A], value_ext [B,1], value_int [B,1])."""
        assert attn_aggregation_op in ["mean", "sum"], 'attention_aggregation_op must be one of ["mean", "sum"]'
        rt = self.runtime()
        rt.sync()
        rt.heads.coef = 0.5 if attn_aggregation_op == "mean" else 1.0
        x = _as_device_image(state, rt.device)
        anchor = next(iter(rt.params.values()))
        if torch.is_grad_enabled():
            return _ActorCriticFn.apply(anchor, rt, x)
        pol, ve, vi = rt.ac_forward(x, x.shape[0])
        return pol.clone(), ve.clone().unsqueeze(1), vi.clone().unsqueeze(1)


class _RNDFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, rt: Runtime, obs):
        B = obs.shape[0]
        ctx.gen = rt._bump_gen("rnd", B)
        pred = rt.rnd_pred.forward(obs, B).clone()
        tgt = rt.rnd_tgt.forward(obs, B).clone()
        ctx.rt, ctx.B = rt, B
        return pred, tgt

    @staticmethod
    def backward(ctx, dpred, dtgt):
        if dpred is not None:
            ctx.rt.check_gen("rnd", ctx.B, ctx.gen, getattr(ctx.rt, "_epoch_gen", 0))
            ctx.rt.rnd_pred.backward(dpred.contiguous().to(torch.bfloat16))
        return None, None, None


class RNDModel(nn.Module):
    """model.py:357-461, ``original_RND`` branch (the only one the ViT configs use)."""

    def __init__(self, input_size=32, output_size=512, train_method="modified_RND"):
        super().__init__()
        assert train_method == "original_RND", "hot path covers TrainMethod = original_RND (SURVEY fact 7)"
        self.input_size, self.output_size = input_size, output_size
        feature_output = 7 * 7 * 64

        def tower():
            return [nn.Conv2d(1, 32, kernel_size=8, stride=4), nn.LeakyReLU(), nn.Conv2d(32, 64, kernel_size=4, stride=2),
                    nn.LeakyReLU(), nn.Conv2d(64, 64, kernel_size=3, stride=1), nn.LeakyReLU(), Flatten(),
                    nn.Linear(feature_output, output_size)]
        self.predictor = nn.Sequential(*tower(), nn.ReLU(), nn.Linear(output_size, output_size), nn.ReLU(),
                                       nn.Linear(output_size, output_size))
        self.target = nn.Sequential(*tower())
        for p in self.modules():                                   # model.py:445-451
            if isinstance(p, (nn.Conv2d, nn.Linear)):
                init.orthogonal_(p.weight, np.sqrt(2))
                p.bias.data.zero_()
        for param in self.target.parameters():                     # model.py:453-455
            param.requires_grad = False
        self._rt = None
        self._agent_rt = None

    def runtime(self) -> Runtime:
        if self._agent_rt is not None:
            return self._agent_rt()
        return _runtime_for(self, "rnd.")

    def forward(self, next_obs):
        """model.py:457-461 -> (predict_feature, target_feature), fp32 [B, 512]."""
        rt = self.runtime()
        rt.sync()
        x = next_obs if torch.is_tensor(next_obs) else torch.as_tensor(next_obs)
        x = x.to(rt.device, dtype=torch.float32).contiguous()
        anchor = next(iter(rt.params.values()))
        if torch.is_grad_enabled():
            return _RNDFn.apply(anchor, rt, x)
        B = x.shape[0]
        rt._bump_gen("rnd", B)
        return rt.rnd_pred.forward(x, B).clone(), rt.rnd_tgt.forward(x, B).clone()
