"""Importable alias for the package directory (whose mandated name contains hyphens)."""
import importlib
import sys

_PKG = "explorative-attention-vit-for-model-predictive-exploration-in-reinforcement-learning_b200"
_mod = importlib.import_module(_PKG)
sys.modules[__name__] = _mod
