"""CPU oracle: a plain torch-fp32 / numpy restatement of the reference learner hot path.

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / ``--impl reference`` legs.  The product package must never import it.

Pinning: the reference ships no golden vectors or tests for this path, so
the oracle is pinned against outputs of the reference itself, run in the build container by
``tests/golden/make_golden.py`` (imports /root/reference) and committed under ``tests/golden/``.
``tests/test_oracle_golden.py`` re-checks the oracle against those fixtures on every CPU run.

Every function cites the reference file:line (relative to /root/reference) it restates.
Parameters live in a flat ``dict[str, Tensor]`` keyed by the reference's ``state_dict`` names
(``model.*`` for CnnActorCriticNetwork, ``rnd.*`` for RNDModel) so that the same dict can be
``load_state_dict``-ed into the reference classes.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------
# configuration
# ----------------------------------------------------------------------------------------------
@dataclass
class OracleConfig:
    """The subset of .conf keys the hot path reads ("Config / flags")."""

    impl: str = "lucidrains"          # ViT_implementation_type: 0 -> "lucidrains", 1 -> "hg"; "cnn" = the original RND
                                      # CNN backbone the reference keeps as commented-out code (model.py:110-178)
    image: int = 84                   # PreProcHeight
    channels: int = 4                 # StateStackSize
    patch: int = 6                    # ViTlucidrains_patch_size / ViTHG_patch_size
    dim: int = 256                    # ViTlucidrains_dim / ViTHG_hidden_size
    depth: int = 3                    # ViTlucidrains_depth / ViTHG_num_hidden_layers
    heads: int = 8                    # ViTlucidrains_heads / ViTHG_num_attention_heads
    dim_head: int = 32                # ViTlucidrains_dim_head (HG: hidden/heads)
    mlp_dim: int = 1024               # ViTlucidrains_mlp_dim / ViTHG_intermediate_size
    use_explorative: bool = True      # ViT*_use_explorativeAttn
    ln_eps: float = 1e-5              # nn.LayerNorm default (HG: ViTHG_layer_norm_eps)
    n_actions: int = 18
    rnd_out: int = 512
    # PPO / RND hyper-parameters (train.py:101-123)
    gamma: float = 0.999
    int_gamma: float = 0.99
    lam: float = 0.95
    lr: float = 1e-4
    ent_coef: float = 0.001
    ppo_eps: float = 0.1
    ext_coef: float = 2.0
    int_coef: float = 1.0
    update_proportion: float = 0.25   # agents.py:46 default (UpdateProportion is never read)
    epoch: int = 4
    mini_batch: int = 32

    @property
    def n_patches(self) -> int:
        return (self.image // self.patch) ** 2

    @property
    def patch_dim(self) -> int:
        return self.channels * self.patch * self.patch


def param_shapes(cfg: OracleConfig) -> Dict[str, Tuple[int, ...]]:
    """Reference ``RNDAgent.state_dict()`` names -> shapes (probed)."""
    s: Dict[str, Tuple[int, ...]] = {}
    D, A = cfg.dim, cfg.n_actions
    if cfg.impl == "lucidrains":
        inner = cfg.heads * cfg.dim_head
        f = "model.feature."
        s[f + "pos_embedding"] = (1, cfg.n_patches + 1, D)                       # vit.py:116
        if cfg.use_explorative:
            s[f + "exploration_token"] = (1, 1, D)                               # vit.py:119
            s[f + "exploitation_token"] = (1, 1, D)                              # vit.py:120
        else:
            s[f + "cls_token"] = (1, 1, D)                                       # vit.py:122
        s[f + "to_patch_embedding.1.weight"] = (cfg.patch_dim,)                  # vit.py:111
        s[f + "to_patch_embedding.1.bias"] = (cfg.patch_dim,)
        s[f + "to_patch_embedding.2.weight"] = (D, cfg.patch_dim)                # vit.py:112
        s[f + "to_patch_embedding.2.bias"] = (D,)
        s[f + "to_patch_embedding.3.weight"] = (D,)                              # vit.py:113
        s[f + "to_patch_embedding.3.bias"] = (D,)
        s[f + "transformer.norm.weight"] = (D,)                                  # vit.py:78
        s[f + "transformer.norm.bias"] = (D,)
        for i in range(cfg.depth):
            a = f + f"transformer.layers.{i}.0."
            s[a + "norm.weight"] = (D,)                                          # vit.py:47
            s[a + "norm.bias"] = (D,)
            s[a + "to_qkv.weight"] = (3 * inner, D)                              # vit.py:52 (no bias)
            s[a + "to_out.0.weight"] = (D, inner)                                # vit.py:55
            s[a + "to_out.0.bias"] = (D,)
            m = f + f"transformer.layers.{i}.1.net."
            s[m + "0.weight"] = (D,)                                             # vit.py:28
            s[m + "0.bias"] = (D,)
            s[m + "1.weight"] = (cfg.mlp_dim, D)                                 # vit.py:29
            s[m + "1.bias"] = (cfg.mlp_dim,)
            s[m + "4.weight"] = (D, cfg.mlp_dim)                                 # vit.py:32
            s[m + "4.bias"] = (D,)
    elif cfg.impl == "hg":
        f = "model.feature."
        e = f + "embeddings."
        s[e + "exploration_token"] = (1, 1, D)                                   # vit_hg.py:56
        s[e + "exploitation_token"] = (1, 1, D)                                  # vit_hg.py:57
        s[e + "patch_embeddings.projection.weight"] = (D, cfg.channels, cfg.patch, cfg.patch)
        s[e + "patch_embeddings.projection.bias"] = (D,)
        s[e + "position_embeddings"] = (1, cfg.n_patches + 1, D)                 # vit_hg.py:62
        for i in range(cfg.depth):
            l = f + f"encoder.layer.{i}."
            for nm in ("query", "key", "value"):
                s[l + f"attention.attention.{nm}.weight"] = (D, D)
                s[l + f"attention.attention.{nm}.bias"] = (D,)
            s[l + "attention.output.dense.weight"] = (D, D)
            s[l + "attention.output.dense.bias"] = (D,)
            s[l + "intermediate.dense.weight"] = (cfg.mlp_dim, D)
            s[l + "intermediate.dense.bias"] = (cfg.mlp_dim,)
            s[l + "output.dense.weight"] = (D, cfg.mlp_dim)
            s[l + "output.dense.bias"] = (D,)
            s[l + "layernorm_before.weight"] = (D,)
            s[l + "layernorm_before.bias"] = (D,)
            s[l + "layernorm_after.weight"] = (D,)
            s[l + "layernorm_after.bias"] = (D,)
        s[f + "layernorm.weight"] = (D,)
        s[f + "layernorm.bias"] = (D,)
        s[f + "pooler.dense.weight"] = (D, D)
        s[f + "pooler.dense.bias"] = (D,)
    elif cfg.impl == "cnn":
        # model.py:110-135 (commented out upstream; BASELINE configs[1], SURVEY 8c "CNN-backbone config"): nn.Sequential
        # indices 0 conv8s4, 2 conv4s2, 4 conv3s1, 6 Flatten, 7 Linear(3136, 256), 9 Linear(256, dim = 448)
        f = "model.feature."
        s[f + "0.weight"] = (32, cfg.channels, 8, 8)
        s[f + "0.bias"] = (32,)
        s[f + "2.weight"] = (64, 32, 4, 4)
        s[f + "2.bias"] = (64,)
        s[f + "4.weight"] = (64, 64, 3, 3)
        s[f + "4.bias"] = (64,)
        s[f + "7.weight"] = (256, 7 * 7 * 64)
        s[f + "7.bias"] = (256,)
        s[f + "9.weight"] = (D, 256)
        s[f + "9.bias"] = (D,)
    else:
        raise ValueError(cfg.impl)
    # heads, model.py:227-246
    s["model.actor.0.weight"] = (D, D)
    s["model.actor.0.bias"] = (D,)
    s["model.actor.2.weight"] = (A, D)
    s["model.actor.2.bias"] = (A,)
    s["model.extra_layer.0.weight"] = (D, D)
    s["model.extra_layer.0.bias"] = (D,)
    s["model.critic_ext.weight"] = (1, D)
    s["model.critic_ext.bias"] = (1,)
    s["model.critic_int.weight"] = (1, D)
    s["model.critic_int.bias"] = (1,)
    # RND original_RND branch, model.py:366-416
    R = cfg.rnd_out
    for net, fcs in (("predictor", (7, 9, 11)), ("target", (7,))):
        r = f"rnd.{net}."
        s[r + "0.weight"] = (32, 1, 8, 8)
        s[r + "0.bias"] = (32,)
        s[r + "2.weight"] = (64, 32, 4, 4)
        s[r + "2.bias"] = (64,)
        s[r + "4.weight"] = (64, 64, 3, 3)
        s[r + "4.bias"] = (64,)
        for j, k in enumerate(fcs):
            s[r + f"{k}.weight"] = (R, 7 * 7 * 64 if j == 0 else R)
            s[r + f"{k}.bias"] = (R,)
    return s


def init_params(cfg: OracleConfig, seed: int = 0, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Deterministic, torch-RNG-independent weights (numpy PCG64) with reference-like scales.

    The reference initialises through torch's global RNG (model.py:249-263, :445-451); goldens
    must not depend on that stream, so the fixture generator loads THESE weights into the
    reference modules (``load_state_dict(strict=True)``), which also pins names and shapes.
    """
    rng = np.random.default_rng(seed)
    out: Dict[str, torch.Tensor] = {}
    for name, shape in param_shapes(cfg).items():
        leaf = name.split(".")[-1]
        is_norm = ("norm" in name) or ("to_patch_embedding.1" in name) or ("to_patch_embedding.3" in name) \
            or (".net.0." in name)
        if "token" in name or "pos_embedding" in name or "position_embeddings" in name:
            a = rng.standard_normal(shape) * (0.02 if cfg.impl == "hg" else 1.0)
        elif is_norm:
            a = (1.0 + 0.1 * rng.standard_normal(shape)) if leaf == "weight" else 0.05 * rng.standard_normal(shape)
        elif leaf == "bias":
            a = 0.02 * rng.standard_normal(shape)
        else:
            fan_in = int(np.prod(shape[1:]))
            gain = 1.0
            if name.startswith("model.critic") or name.startswith("model.actor"):
                gain = 0.1          # reference uses orthogonal gain 0.01 (model.py:249-258)
            elif name.startswith("model.extra_layer"):
                gain = 0.3          # reference: orthogonal gain 0.1 (model.py:260-263)
            elif name.startswith("rnd.") or (cfg.impl == "cnn" and name.startswith("model.feature.")):
                gain = math.sqrt(2.0)   # model.py:447,451 (RND towers), model.py:149-156 (CNN backbone)
            a = rng.standard_normal(shape) * (gain / math.sqrt(fan_in))
        out[name] = torch.tensor(a, dtype=dtype)
    return out


# ----------------------------------------------------------------------------------------------
# lucidrains ViT with explorative attention  (vit.py)
# ----------------------------------------------------------------------------------------------
EXPLORATIVE, EXPLOITATIVE, CLS = 0, 1, 2   # vit.py:14-17 ViT_Attn values

# nn.Dropout sites of the reference (vit.py:31,33,45,56,158; HF hidden / attention-probs dropout).  torch's Philox stream
# cannot be reproduced by another implementation, so parity with dropout ON is checked with EXPLICIT masks: a test installs
# DROPOUT_HOOK(kind, layer, x, pass_id) -> x * mask, fed with the very masks the CUDA kernels generate.  None = identity
# (dropout 0, the configuration every golden fixture uses).
DROPOUT_HOOK = None


def _drop(kind: str, layer: int, x: torch.Tensor, pass_id) -> torch.Tensor:
    return x if DROPOUT_HOOK is None else DROPOUT_HOOK(kind, layer, x, pass_id)


def patchify_lucid(img: torch.Tensor, p: int) -> torch.Tensor:
    """vit.py:110 -- 'b c (h p1) (w p2) -> b (h w) (p1 p2 c)' (channel fastest)."""
    b, c, H, W = img.shape
    x = img.reshape(b, c, H // p, p, W // p, p)
    return x.permute(0, 2, 4, 3, 5, 1).reshape(b, (H // p) * (W // p), p * p * c)


def lucid_embed(P, img, attn_type: int, cfg: OracleConfig, pre="model.feature.") -> torch.Tensor:
    """vit.py:138-158 (dropout = identity in parity runs).

    Reproduces the token-prepend bug: both branches at vit.py:142 and :146 test
    EXPLOITATIVE_ATTN, so EXPLORATIVE adds no token and no positional embedding, and
    EXPLOITATIVE prepends ``exploration_token``.
    """
    x = patchify_lucid(img, cfg.patch)
    x = F.layer_norm(x, (cfg.patch_dim,), P[pre + "to_patch_embedding.1.weight"], P[pre + "to_patch_embedding.1.bias"], 1e-5)
    x = F.linear(x, P[pre + "to_patch_embedding.2.weight"], P[pre + "to_patch_embedding.2.bias"])
    x = F.layer_norm(x, (cfg.dim,), P[pre + "to_patch_embedding.3.weight"], P[pre + "to_patch_embedding.3.bias"], 1e-5)
    b, n, _ = x.shape
    if cfg.use_explorative:
        if attn_type == EXPLOITATIVE:
            tok = P[pre + "exploration_token"].expand(b, 1, cfg.dim)
            x = torch.cat((tok, x), dim=1) + P[pre + "pos_embedding"][:, : n + 1]
        elif attn_type != EXPLORATIVE:
            raise AssertionError("explorative ViT takes EXPLORATIVE/EXPLOITATIVE")
    else:
        if attn_type != CLS:
            raise Exception("Must use attn_type=CLS")                            # vit.py:153
        tok = P[pre + "cls_token"].expand(b, 1, cfg.dim)
        x = torch.cat((tok, x), dim=1) + P[pre + "pos_embedding"][:, : n + 1]
    return _drop("emb", 0, x, attn_type)                                         # vit.py:158


def lucid_attention(P, x, i: int, cfg: OracleConfig, pre="model.feature.", pass_id=None) -> torch.Tensor:
    """vit.py:60-73."""
    a = pre + f"transformer.layers.{i}.0."
    h, d = cfg.heads, cfg.dim_head
    b, n, _ = x.shape
    y = F.layer_norm(x, (cfg.dim,), P[a + "norm.weight"], P[a + "norm.bias"], 1e-5)
    qkv = F.linear(y, P[a + "to_qkv.weight"])
    q, k, v = (t.reshape(b, n, h, d).transpose(1, 2) for t in qkv.chunk(3, dim=-1))
    dots = torch.matmul(q, k.transpose(-1, -2)) * (d ** -0.5)
    attn = _drop("attn_p", i, dots.softmax(dim=-1), pass_id)                     # vit.py:69-70
    out = torch.matmul(attn, v).transpose(1, 2).reshape(b, n, h * d)
    return _drop("attn_out", i, F.linear(out, P[a + "to_out.0.weight"], P[a + "to_out.0.bias"]), pass_id)   # vit.py:54-57


def lucid_ff(P, x, i: int, cfg: OracleConfig, pre="model.feature.", pass_id=None) -> torch.Tensor:
    """vit.py:27-37 (nn.GELU default = exact erf)."""
    m = pre + f"transformer.layers.{i}.1.net."
    y = F.layer_norm(x, (cfg.dim,), P[m + "0.weight"], P[m + "0.bias"], 1e-5)
    y = _drop("act", i, F.gelu(F.linear(y, P[m + "1.weight"], P[m + "1.bias"])), pass_id)                  # vit.py:30-31
    return _drop("ff_out", i, F.linear(y, P[m + "4.weight"], P[m + "4.bias"]), pass_id)                    # vit.py:32-33


def lucid_vit(P, img, attn_type: int, cfg: OracleConfig, pre="model.feature.") -> torch.Tensor:
    """vit.py:136-167 with num_classes = -1, pool = 'cls'."""
    x = lucid_embed(P, img, attn_type, cfg, pre)
    for i in range(cfg.depth):                                                   # vit.py:86-89
        x = lucid_attention(P, x, i, cfg, pre, attn_type) + x
        x = lucid_ff(P, x, i, cfg, pre, attn_type) + x
    x = F.layer_norm(x, (cfg.dim,), P[pre + "transformer.norm.weight"], P[pre + "transformer.norm.bias"], 1e-5)
    return x[:, 0]                                                               # vit.py:162


# ----------------------------------------------------------------------------------------------
# HuggingFace-style ViT with explorative attention (vit_hg.py + transformers ViT arithmetic)
# ----------------------------------------------------------------------------------------------
def hg_vit(P, img, cfg: OracleConfig, pre="model.feature.") -> Tuple[torch.Tensor, torch.Tensor]:
    """vit_hg.py:101-163 (embeddings), :316-355 (two encoder passes + final LN).

    Encoder arithmetic follows transformers' ``ViTLayer``/``ViTSelfAttention`` (pinned 4.37.0,
    not vendored in the reference): pre-LN, separate q/k/v Linear with bias, softmax(QK^T/sqrt(d)),
    dense, residual, LN, dense+GELU(erf), dense, residual.  Returns the two [B,S,D] sequences.
    """
    e = pre + "embeddings."
    D, h = cfg.dim, cfg.heads
    d = D // h
    x = F.conv2d(img, P[e + "patch_embeddings.projection.weight"], P[e + "patch_embeddings.projection.bias"], stride=cfg.patch)
    x = x.flatten(2).transpose(1, 2)                                             # [B, n, D]
    b = x.shape[0]
    outs = []
    for pass_id, tok in enumerate(("exploration_token", "exploitation_token")):  # vit_hg.py:121-145
        s = torch.cat((P[e + tok].expand(b, -1, -1), x), dim=1) + P[e + "position_embeddings"]
        s = _drop("emb", 0, s, pass_id)                                          # ViTEmbeddings dropout (hidden_dropout_prob)
        for i in range(cfg.depth):
            l = pre + f"encoder.layer.{i}."
            y = F.layer_norm(s, (D,), P[l + "layernorm_before.weight"], P[l + "layernorm_before.bias"], cfg.ln_eps)
            n = y.shape[1]
            q = F.linear(y, P[l + "attention.attention.query.weight"], P[l + "attention.attention.query.bias"])
            k = F.linear(y, P[l + "attention.attention.key.weight"], P[l + "attention.attention.key.bias"])
            v = F.linear(y, P[l + "attention.attention.value.weight"], P[l + "attention.attention.value.bias"])
            q, k, v = (t.reshape(b, n, h, d).transpose(1, 2) for t in (q, k, v))
            pr = _drop("attn_p", i, (torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(d)).softmax(dim=-1), pass_id)
            o = torch.matmul(pr, v).transpose(1, 2).reshape(b, n, D)
            s = _drop("attn_out", i, F.linear(o, P[l + "attention.output.dense.weight"], P[l + "attention.output.dense.bias"]), pass_id) + s
            y = F.layer_norm(s, (D,), P[l + "layernorm_after.weight"], P[l + "layernorm_after.bias"], cfg.ln_eps)
            y = F.gelu(F.linear(y, P[l + "intermediate.dense.weight"], P[l + "intermediate.dense.bias"]))
            s = _drop("ff_out", i, F.linear(y, P[l + "output.dense.weight"], P[l + "output.dense.bias"]), pass_id) + s
        outs.append(F.layer_norm(s, (D,), P[pre + "layernorm.weight"], P[pre + "layernorm.bias"], cfg.ln_eps))
    return outs[0], outs[1]


# ----------------------------------------------------------------------------------------------
# actor-critic wrapper and RND nets (model.py)
# ----------------------------------------------------------------------------------------------
def _heads(P, x_int_feat, x_ext_feat, x_pol_feat, ext_critic: str):
    def value(feat, critic):
        e = F.relu(F.linear(feat, P["model.extra_layer.0.weight"], P["model.extra_layer.0.bias"]))
        return F.linear(e + feat, P[f"model.{critic}.weight"], P[f"model.{critic}.bias"])
    value_int = value(x_int_feat, "critic_int")
    value_ext = value(x_ext_feat, ext_critic)
    a = F.relu(F.linear(x_pol_feat, P["model.actor.0.weight"], P["model.actor.0.bias"]))
    policy = F.linear(a, P["model.actor.2.weight"], P["model.actor.2.bias"])
    return policy, value_ext, value_int


def actor_critic_forward(P, state, cfg: OracleConfig, attn_aggregation_op: str = "mean"):
    """model.py:266-352, DISCRETE action space.  Returns (policy, value_ext, value_int).

    lucidrains/explorative: model.py:272-296.  lucidrains/CLS: model.py:298-304.
    HG: model.py:310-336 incl. the head bug: value_ext uses ``critic_int``.
    """
    if cfg.impl == "cnn":
        # model.py:110-135 backbone; forward as in the upstream RND agent the comment block was taken from (README.md:168,
        # SURVEY 8c): policy = actor(x), value = critic(extra_layer(x) + x) for both critics
        x = cnn_backbone(P, state)
        return _heads(P, x, x, x, "critic_ext")
    if cfg.impl == "lucidrains":
        if cfg.use_explorative:
            xe = lucid_vit(P, state, EXPLORATIVE, cfg)
            xx = lucid_vit(P, state, EXPLOITATIVE, cfg)
            comb = torch.stack((xe, xx), dim=1).mean(dim=1) if attn_aggregation_op == "mean" else xe + xx
            return _heads(P, xe, xx, comb, "critic_ext")
        xc = lucid_vit(P, state, CLS, cfg)
        return _heads(P, xc, xc, xc, "critic_ext")
    s_e, s_x = hg_vit(P, state, cfg)
    xe, xx = s_e[:, 0, :], s_x[:, 0, :]                                          # model.py:316,320
    comb = torch.stack((xe, xx), dim=1).mean(dim=1) if attn_aggregation_op == "mean" else xe + xx
    return _heads(P, xe, xx, comb, "critic_int")                                 # model.py:321 (bug kept)


def cnn_backbone(P, state, pre="model.feature.") -> torch.Tensor:
    """model.py:110-135 (commented out upstream): conv8s4-ReLU-conv4s2-ReLU-conv3s1-ReLU-Flatten-Linear(3136,256)-ReLU-
    Linear(256,448)-ReLU on the stacked frames [B, StateStackSize, 84, 84]."""
    x = F.relu(F.conv2d(state, P[pre + "0.weight"], P[pre + "0.bias"], stride=4))
    x = F.relu(F.conv2d(x, P[pre + "2.weight"], P[pre + "2.bias"], stride=2))
    x = F.relu(F.conv2d(x, P[pre + "4.weight"], P[pre + "4.bias"], stride=1))
    x = F.relu(F.linear(x.flatten(1), P[pre + "7.weight"], P[pre + "7.bias"]))
    return F.relu(F.linear(x, P[pre + "9.weight"], P[pre + "9.bias"]))


def rnd_net(P, obs, net: str) -> torch.Tensor:
    """model.py:368-416: conv8s4-LReLU-conv4s2-LReLU-conv3s1-LReLU-flatten-FC(512)[-ReLU-FC-ReLU-FC]."""
    r = f"rnd.{net}."
    x = F.leaky_relu(F.conv2d(obs, P[r + "0.weight"], P[r + "0.bias"], stride=4))
    x = F.leaky_relu(F.conv2d(x, P[r + "2.weight"], P[r + "2.bias"], stride=2))
    x = F.leaky_relu(F.conv2d(x, P[r + "4.weight"], P[r + "4.bias"], stride=1))
    x = F.linear(x.flatten(1), P[r + "7.weight"], P[r + "7.bias"])
    if net == "predictor":
        x = F.linear(F.relu(x), P[r + "9.weight"], P[r + "9.bias"])
        x = F.linear(F.relu(x), P[r + "11.weight"], P[r + "11.bias"])
    return x


def intrinsic_reward(P, next_obs: np.ndarray) -> np.ndarray:
    """agents.py:210-218: (target - predictor)^2 .mean(1); float64 numpy in, float32 numpy out."""
    with torch.no_grad():
        x = torch.FloatTensor(next_obs)
        return (rnd_net(P, x, "target") - rnd_net(P, x, "predictor")).pow(2).mean(1).numpy()


def get_action(P, state: np.ndarray, cfg: OracleConfig, u: Optional[np.ndarray] = None):
    """agents.py:187-208.  ``u`` = the np.random.rand(E) draw (passed in to stay deterministic)."""
    with torch.no_grad():
        policy, v_ext, v_int = actor_critic_forward(P, torch.Tensor(state).float(), cfg)
        prob = F.softmax(policy, dim=-1).numpy()
    if u is None:
        u = np.random.rand(prob.shape[0])
    action = (prob.cumsum(axis=1) > np.expand_dims(u, 1)).argmax(axis=1)         # agents.py:206-208
    return action, v_ext.numpy().squeeze(), v_int.numpy().squeeze(), policy.numpy()


# ----------------------------------------------------------------------------------------------
# loss and update (agents.py:263-535)
# ----------------------------------------------------------------------------------------------
def ppo_rnd_loss(P, cfg: OracleConfig, s_batch, target_ext, target_int, y, adv, next_obs, old_logits, mask):
    """One minibatch loss, agents.py:333-338 (RND) and :455-493 (PPO).  ``mask`` is the float 0/1
    Bernoulli mask the reference draws with ``torch.rand(B) < update_proportion`` (agents.py:336-337).
    Returns (loss, dict of scalar terms, (policy, value_ext, value_int))."""
    log_prob_old = torch.log_softmax(old_logits, dim=-1).gather(1, y[:, None]).squeeze(1)   # :302-303
    pred, tgt = rnd_net(P, next_obs, "predictor"), rnd_net(P, next_obs, "target")
    per = (pred - tgt.detach()).pow(2).mean(-1)                                             # :335
    rnd_loss = (per * mask).sum() / torch.max(mask.sum(), torch.ones(()))                   # :338
    policy, v_ext, v_int = actor_critic_forward(P, s_batch, cfg)                            # :455
    logp_all = torch.log_softmax(policy, dim=-1)
    log_prob = logp_all.gather(1, y[:, None]).squeeze(1)                                    # :457
    ratio = torch.exp(log_prob - log_prob_old)                                              # :466
    surr1 = ratio * adv
    surr2 = torch.clamp(ratio, 1.0 - cfg.ppo_eps, 1.0 + cfg.ppo_eps) * adv                  # :469-472
    actor_loss = -torch.min(surr1, surr2).mean()                                            # :474
    critic_ext = F.mse_loss(v_ext.sum(1), target_ext)                                       # :476
    critic_int = F.mse_loss(v_int.sum(1), target_int)                                       # :479
    entropy = -(logp_all.exp() * logp_all).sum(-1).mean()                                   # :483
    loss = actor_loss + 0.5 * (critic_ext + critic_int) - cfg.ent_coef * entropy + rnd_loss  # :493
    terms = dict(loss=loss, actor=actor_loss, critic_ext=critic_ext, critic_int=critic_int,
                 entropy=entropy, rnd=rnd_loss)
    return loss, {k: float(v.detach()) for k, v in terms.items()}, (policy, v_ext, v_int)


def ppo_rnd_backward_chunked(P, cfg: OracleConfig, s_batch, target_ext, target_int, y, adv, next_obs, old_logits, mask,
                             chunk: int = 128):
    """The same minibatch loss and gradients as ``ppo_rnd_loss(...)[0].backward()``, evaluated ``chunk`` samples at a time
    (a 512-sample minibatch of the cfg3 model keeps ~10 GB of fp32 autograd state otherwise).  Every PPO term is a mean
    over the minibatch (agents.py:474-483) and the RND term a masked sum over a minibatch-wide count (agents.py:338), so
    the chunk losses recombine exactly: PPO terms weighted n_c / B, the RND term by max(sum mask_c, 1) / max(sum mask, 1).
    The actor-critic and the RND predictor share no parameter, so each part is differentiated with respect to its own
    tensors only.  Accumulates into ``P[k].grad``; returns the dict of scalar terms."""
    B = len(s_batch)
    msum = max(float(mask.sum()), 1.0)
    ac_params = [P[k] for k in P if k.startswith("model.") and P[k].requires_grad]
    rnd_params = [P[k] for k in P if k.startswith("rnd.predictor.") and P[k].requires_grad]
    acc = dict(actor=0.0, critic_ext=0.0, critic_int=0.0, entropy=0.0, rnd=0.0)
    for lo in range(0, B, chunk):
        sl = slice(lo, min(B, lo + chunk))
        w_ppo = (sl.stop - sl.start) / B
        w_rnd = max(float(mask[sl].sum()), 1.0) / msum
        loss, terms, _ = ppo_rnd_loss(P, cfg, s_batch[sl], target_ext[sl], target_int[sl], y[sl], adv[sl], next_obs[sl],
                                      old_logits[sl], mask[sl])
        # one graph, two differently weighted parts: d loss / d(actor-critic) carries only the PPO terms, d loss / d(predictor)
        # only the RND term
        g_ac = torch.autograd.grad(loss, ac_params, retain_graph=True, allow_unused=True)
        g_rnd = torch.autograd.grad(loss, rnd_params, allow_unused=True)
        for p_, g in zip(ac_params, g_ac):
            if g is not None:
                p_.grad = g * w_ppo if p_.grad is None else p_.grad + g * w_ppo
        for p_, g in zip(rnd_params, g_rnd):
            if g is not None:
                p_.grad = g * w_rnd if p_.grad is None else p_.grad + g * w_rnd
        for k in ("actor", "critic_ext", "critic_int", "entropy"):
            acc[k] += terms[k] * w_ppo
        acc["rnd"] += terms["rnd"] * w_rnd
    acc["loss"] = acc["actor"] + 0.5 * (acc["critic_ext"] + acc["critic_int"]) - cfg.ent_coef * acc["entropy"] + acc["rnd"]
    return acc


def trainable_names(P) -> list:
    """agents.py:141-164: model.* and rnd.predictor.* (rnd.target is frozen, model.py:453-455)."""
    return [k for k in P if k.startswith("model.") or k.startswith("rnd.predictor.")]


def train_model(P, cfg: OracleConfig, states, target_ext, target_int, y, adv, next_obs_norm, old_policy,
                n_shards: int = 1, max_steps: Optional[int] = None, opt=None):
    """agents.py:263-535 restated (DISCRETE, original_RND, no SSL, no grad clipping).

    ``states`` float32 [N,C,H,W] (already /255), ``target_*``/``adv`` float64 [N], ``y`` int64 [N],
    ``next_obs_norm`` float64 [N,1,H,W], ``old_policy`` float32 [T,E,A] (step-major).
    Consumes ``np.random.shuffle`` and ``torch.rand`` on the global generators exactly like the
    reference (agents.py:276, :336).  ``n_shards`` > 1 is the multi-GPU oracle of SURVEY 8(c/e):
    each contiguous env shard runs the reference minibatch on its own slice and the gradients are
    averaged before one Adam step.  Mutates ``P`` in place; returns the list of per-step terms.
    """
    N = len(states)
    batch = N // cfg.mini_batch
    names = trainable_names(P)
    for k in names:
        P[k].requires_grad_(True)
    if opt is None:
        opt = torch.optim.Adam([P[k] for k in names], lr=cfg.lr)                # agents.py:129
    A = old_policy.shape[-1]
    old_flat = torch.tensor(old_policy).permute(1, 0, 2).contiguous().view(-1, A)           # :301
    t_states = torch.FloatTensor(states)
    t_ext, t_int, t_adv = torch.FloatTensor(target_ext), torch.FloatTensor(target_int), torch.FloatTensor(adv)
    t_y = torch.LongTensor(y)
    t_obs = torch.FloatTensor(next_obs_norm)
    log = []
    shard = N // n_shards
    lbatch = batch // n_shards
    # every rank holds the same numpy / torch seed (train.py:52): one permutation and one mask
    # stream, reused by every shard.  The permutation array persists across epochs (agents.py:270).
    perm = np.arange(shard)
    for _ in range(cfg.epoch):
        np.random.shuffle(perm)                                                  # agents.py:276
        for j in range(shard // lbatch):                                         # agents.py:284
            opt.zero_grad()
            terms_acc = None
            mask = (torch.rand(lbatch) < cfg.update_proportion).float()          # agents.py:336-337
            for r in range(n_shards):
                idx = torch.from_numpy(perm[lbatch * j: lbatch * (j + 1)] + r * shard)
                loss, terms, _ = ppo_rnd_loss(P, cfg, t_states[idx], t_ext[idx], t_int[idx], t_y[idx],
                                              t_adv[idx], t_obs[idx], old_flat[idx], mask)
                (loss / n_shards).backward()
                terms_acc = terms if terms_acc is None else {k: terms_acc[k] + terms[k] for k in terms}
            opt.step()                                                           # agents.py:508
            log.append({k: v / n_shards for k, v in terms_acc.items()})
            if max_steps is not None and len(log) >= max_steps:
                for k in names:
                    P[k].requires_grad_(False)
                return log
    for k in names:
        P[k].requires_grad_(False)
    return log


# ----------------------------------------------------------------------------------------------
# numpy numerics (utils.py, train.py glue)
# ----------------------------------------------------------------------------------------------
def make_train_data(reward, done, value, gamma, num_step, num_worker, lam=0.95, use_gae=True):
    """utils.py:42-67.  numpy promotion is part of the contract: ``gae`` starts
    as an int64 array of shape (1,), ``gamma * value`` is rounded in the dtype of ``value``."""
    discounted_return = np.empty([num_worker, num_step])
    if use_gae:
        gae = np.zeros_like([num_worker, ])
        for t in range(num_step - 1, -1, -1):
            delta = reward[:, t] + gamma * value[:, t + 1] * (1 - done[:, t]) - value[:, t]
            gae = delta + gamma * lam * (1 - done[:, t]) * gae
            discounted_return[:, t] = gae + value[:, t]
        adv = discounted_return - value[:, :-1]
    else:
        running_add = value[:, -1]
        for t in range(num_step - 1, -1, -1):
            running_add = reward[:, t] + gamma * running_add * (1 - done[:, t])
            discounted_return[:, t] = running_add
        adv = discounted_return - value[:, :-1]
    return discounted_return.reshape([-1]), adv.reshape([-1])


class RunningMeanStd:
    """utils.py:70-115 (original_RND / reward_rms branch): Chan parallel-variance merge, float64."""

    def __init__(self, epsilon=1e-4, shape=()):
        self.mean = np.zeros(shape, "float64")
        self.var = np.ones(shape, "float64")
        self.count = epsilon

    def update(self, x):
        self.update_from_moments(np.mean(x, axis=0), np.var(x, axis=0), x.shape[0])

    def update_from_moments(self, batch_mean, batch_var, batch_count):
        delta = batch_mean - self.mean
        tot = self.count + batch_count
        new_mean = self.mean + delta * batch_count / tot
        m2 = self.var * self.count + batch_var * batch_count + np.square(delta) * self.count * batch_count / tot
        self.mean, self.var, self.count = new_mean, m2 / tot, batch_count + self.count


class RewardForwardFilter:
    """utils.py:118-128."""

    def __init__(self, gamma):
        self.rewems = None
        self.gamma = gamma

    def update(self, rews):
        self.rewems = rews if self.rewems is None else self.rewems * self.gamma + rews
        return self.rewems


def normalize_obs(x, rms: RunningMeanStd):
    """train.py:666 / :855 -- ((x - mean) / sqrt(var)).clip(-5, 5) in float64."""
    return ((x - rms.mean) / np.sqrt(rms.var)).clip(-5, 5)


def normalize_int_reward(total_int_reward, filt: RewardForwardFilter, reward_rms: RunningMeanStd):
    """train.py:736-743.  ``total_int_reward`` float32 [E,T]; returns the normalised float32 [E,T].
    Keeps the count = T (not T*E) quirk of train.py:739-740."""
    per_env = np.array([filt.update(r) for r in total_int_reward.T])
    mean, std, count = np.mean(per_env), np.std(per_env), len(per_env)
    reward_rms.update_from_moments(mean, std ** 2, count)
    out = total_int_reward.copy()
    out /= np.sqrt(reward_rms.var)
    return out


def relayout_rollout(num_step, num_env, total_state, total_reward, total_action, total_done, total_next_obs,
                     total_ext_values, total_int_values, total_policy):
    """train.py:707-719: step-major [T*E, ...] buffers -> env-major (flat sample index e*T + t)."""
    T, E = num_step, num_env
    C, H, W = total_state.shape[1:]
    st = total_state.reshape([T, E, C, H, W]).transpose(1, 0, 2, 3, 4).reshape([-1, C, H, W])
    rw = total_reward.reshape([T, E]).transpose().clip(-1, 1)
    ac = total_action.reshape([T, E]).transpose().reshape([-1])
    dn = total_done.reshape([T, E]).transpose().reshape([E, T])
    no = total_next_obs.reshape([T, E, 1, H, W]).transpose([1, 0, 2, 3, 4]).reshape([E * T, 1, H, W])
    ve = total_ext_values.reshape([T + 1, E]).transpose().reshape(E, T + 1)
    vi = total_int_values.reshape([T + 1, E]).transpose().reshape(E, T + 1)
    po = total_policy.reshape([T, E, -1])
    return st, rw, ac, dn, no, ve, vi, po


def prepare_update(cfg: OracleConfig, T, E, roll, obs_rms, reward_rms, filt):
    """train.py:707-779 + :855 glue for one update, from step-major rollout buffers ``roll``
    (dict with the train.py:582-599 names) to the ``train_model`` argument tuple."""
    st, rw, ac, dn, no, ve, vi, po = relayout_rollout(
        T, E, roll["total_state"], roll["total_reward"], roll["total_action"], roll["total_done"],
        roll["total_next_obs"], roll["total_ext_values"], roll["total_int_values"], roll["total_policy"])
    ir = roll["total_int_reward"].reshape([T, E]).transpose().reshape([E, T])
    ir = normalize_int_reward(ir, filt, reward_rms)
    ext_target, ext_adv = make_train_data(rw, dn, ve, cfg.gamma, T, E, cfg.lam)
    int_target, int_adv = make_train_data(ir, np.zeros_like(ir), vi, cfg.int_gamma, T, E, cfg.lam)
    total_adv = int_adv * cfg.int_coef + ext_adv * cfg.ext_coef                   # train.py:767
    obs_rms.update(no)                                                            # train.py:774
    return (np.float32(st) / 255.0, ext_target, int_target, ac, total_adv, normalize_obs(no, obs_rms), po)


# ----------------------------------------------------------------------------------------------
# synthetic inputs
# ----------------------------------------------------------------------------------------------
def synth_rollout(E: int, T: int, seed: int, A: int = 18, image: int = 84, C: int = 4):
    """Step-major synthetic rollout buffers with the dtypes of train.py:582-599."""
    rng = np.random.default_rng(seed)
    N = E * T
    return dict(
        total_state=rng.integers(0, 256, (N, C, image, image), dtype=np.uint8).astype(np.float64),
        total_next_obs=rng.integers(0, 256, (N, 1, image, image), dtype=np.uint8).astype(np.float64),
        total_reward=rng.normal(0, 1, N).astype(np.float64),
        total_action=rng.integers(0, A, N).astype(np.int64),
        total_done=(rng.random(N) < 0.05),
        total_ext_values=rng.normal(0, 1, E * (T + 1)).astype(np.float32),
        total_int_values=rng.normal(0, 1, E * (T + 1)).astype(np.float32),
        total_policy=rng.normal(0, 1, (N, A)).astype(np.float32),
        total_int_reward=rng.random(N).astype(np.float32),
    )
