"""Importable alias for the package directory (whose mandated name contains hyphens).

``import eavit_b200`` / ``from eavit_b200 import agents`` / ``from eavit_b200.vit import ViT_Attn`` all resolve to
the SAME module objects as the hyphenated package (no duplicate classes / enums)."""
import importlib
import sys

_PKG = "explorative-attention-vit-for-model-predictive-exploration-in-reinforcement-learning_b200"
_mod = importlib.import_module(_PKG)
for _sub in ("_lib", "ops", "config", "dist", "utils", "engine", "vit", "vit_hg", "model", "agents", "custom_ops"):
    _m = importlib.import_module(_PKG + "." + _sub)
    sys.modules[__name__ + "." + _sub] = _m
    setattr(_mod, _sub, _m)
sys.modules[__name__] = _mod
