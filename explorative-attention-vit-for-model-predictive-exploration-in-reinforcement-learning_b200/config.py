"""Config contract of the hot path: the reference's ``.conf`` ``[DEFAULT]`` keys.

The reference parses ``sys.argv`` at import time (config.py:14) and exposes a module-global
``default_config``.  Here the same keys are read from a path (or a dict) explicitly; ``default_config``
is still a module global so that ``model.CnnActorCriticNetwork(...)`` can be constructed exactly like
the reference's (no config argument), after ``load_config(path)``.
"""
from __future__ import annotations

import configparser
from dataclasses import dataclass
from typing import Mapping, Optional, Union

_DEFAULTS = {
    # reference: configs/demo_config.conf key names (values = expGlados3 lucidrains config)
    "TrainMethod": "original_RND", "representationLearningMethod": "None", "freeze_shared_backbone": "False",
    "extracted_feature_embedding_dim": "448", "ViT_implementation_type": "0",
    "ViTlucidrains_use_explorativeAttn": "True", "ViTlucidrains_dim": "256", "ViTlucidrains_patch_size": "6",
    "ViTlucidrains_num_classes": "-1", "ViTlucidrains_depth": "3", "ViTlucidrains_heads": "8",
    "ViTlucidrains_mlp_dim": "1024", "ViTlucidrains_dropout": "0.1", "ViTlucidrains_emb_dropout": "0.1",
    "ViTlucidrains_dim_head": "32",
    "ViTHG_use_explorativeAttn": "True", "ViTHG_hidden_size": "1024", "ViTHG_num_hidden_layers": "12",
    "ViTHG_num_attention_heads": "16", "ViTHG_intermediate_size": "3072", "ViTHG_hidden_dropout_prob": "0.0",
    "ViTHG_attention_probs_dropout_prob": "0.0", "ViTHG_initializer_range": "0.02", "ViTHG_layer_norm_eps": "1e-12",
    "ViTHG_patch_size": "12", "ViTHG_qkv_bias": "True", "ViTHG_encoder_stride": "16",
    "ViTHG_PreProcHeight": "84", "ViTHG_StateStackSize": "4",
    "PPOEps": "0.1", "Entropy": "0.001", "NumStep": "128", "MiniBatch": "32", "Epoch": "4", "LearningRate": "0.0001",
    "StateStackSize": "4", "IntGamma": "0.99", "Gamma": "0.999", "ExtCoef": "2", "IntCoef": "1",
    "UpdateProportion": "0.25", "UseGAE": "True", "GAELambda": "0.95", "PreProcHeight": "84", "ProProcWidth": "84",
    "UseNoisyNet": "False", "UseGPU": "True", "UseGradClipping": "False", "MaxGradNorm": "0.5",
    "verbose_logging": "False", "ObsNormStep": "50",
}


class Config(dict):
    """String-valued mapping with configparser's ``getboolean`` (what the reference's SectionProxy offers)."""

    def getboolean(self, key: str, fallback: Optional[bool] = None) -> bool:
        if key not in self:
            if fallback is None:
                raise KeyError(key)
            return fallback
        v = str(self[key]).strip().lower()
        if v in ("1", "yes", "true", "on"):
            return True
        if v in ("0", "no", "false", "off"):
            return False
        raise ValueError(f"Not a boolean: {self[key]}")


default_config = Config(_DEFAULTS)


def load_config(src: Union[str, Mapping, None] = None, **overrides) -> Config:
    """Load a reference ``.conf`` file (or a mapping) into the module-global ``default_config``."""
    default_config.clear()
    default_config.update(_DEFAULTS)
    if isinstance(src, str):
        cp = configparser.ConfigParser()
        cp.optionxform = str
        if not cp.read(src):
            raise FileNotFoundError(src)
        default_config.update({k: v for k, v in cp["DEFAULT"].items()})
    elif src is not None:
        default_config.update({k: str(v) for k, v in src.items()})
    default_config.update({k: str(v) for k, v in overrides.items()})
    return default_config


@dataclass
class HotPathConfig:
    """Typed view of the keys the kernels need."""
    impl: str
    image: int
    channels: int
    patch: int
    dim: int
    depth: int
    heads: int
    dim_head: int
    mlp_dim: int
    use_explorative: bool
    ln_eps: float
    dropout: float            # after to_out / MLP2 (vit.py:33,56; HF hidden_dropout_prob)
    emb_dropout: float        # after the positional add (vit.py:158)
    attn_dropout: float = -1.0   # on the attention probabilities (vit.py:45); < 0: same as `dropout`
    act_dropout: float = -1.0    # after GELU (vit.py:31); < 0: same as `dropout`; HF has none

    def __post_init__(self):
        if self.attn_dropout < 0:
            self.attn_dropout = self.dropout
        if self.act_dropout < 0:
            self.act_dropout = self.dropout

    @property
    def n_patches(self) -> int:
        return (self.image // self.patch) ** 2

    @property
    def patch_dim(self) -> int:
        return self.channels * self.patch * self.patch

    @staticmethod
    def from_config(cfg: Mapping = None) -> "HotPathConfig":
        c = default_config if cfg is None else cfg
        if int(c["ViT_implementation_type"]) == 0:        # model.py:183-196
            return HotPathConfig(
                impl="lucidrains", image=int(c["PreProcHeight"]), channels=int(c["StateStackSize"]),
                patch=int(c["ViTlucidrains_patch_size"]), dim=int(c["ViTlucidrains_dim"]),
                depth=int(c["ViTlucidrains_depth"]), heads=int(c["ViTlucidrains_heads"]),
                dim_head=int(c["ViTlucidrains_dim_head"]), mlp_dim=int(c["ViTlucidrains_mlp_dim"]),
                use_explorative=Config(c).getboolean("ViTlucidrains_use_explorativeAttn"), ln_eps=1e-5,
                dropout=float(c["ViTlucidrains_dropout"]), emb_dropout=float(c["ViTlucidrains_emb_dropout"]))
        hidden, heads = int(c["ViTHG_hidden_size"]), int(c["ViTHG_num_attention_heads"])   # model.py:200-219
        return HotPathConfig(
            impl="hg", image=int(c["ViTHG_PreProcHeight"]), channels=int(c["ViTHG_StateStackSize"]),
            patch=int(c["ViTHG_patch_size"]), dim=hidden, depth=int(c["ViTHG_num_hidden_layers"]), heads=heads,
            dim_head=hidden // heads, mlp_dim=int(c["ViTHG_intermediate_size"]),
            use_explorative=Config(c).getboolean("ViTHG_use_explorativeAttn"), ln_eps=float(c["ViTHG_layer_norm_eps"]),
            dropout=float(c["ViTHG_hidden_dropout_prob"]), emb_dropout=float(c["ViTHG_hidden_dropout_prob"]),
            attn_dropout=float(c["ViTHG_attention_probs_dropout_prob"]), act_dropout=0.0)
