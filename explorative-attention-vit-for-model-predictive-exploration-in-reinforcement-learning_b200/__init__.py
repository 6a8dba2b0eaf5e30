"""B200-native (sm_100a) learner hot path for the Explorative-Attention ViT + RND/PPO agent.

Hand-written CUDA behind a C ABI (``include/eavit_b200.h`` -> ``libeavit_b200.so``) plus a Python
host side that mirrors the reference's operator surface (``vit.ViT``, ``model.CnnActorCriticNetwork``,
``model.RNDModel``, ``agents.RNDAgent``, ``utils.make_train_data`` / ``RunningMeanStd`` /
``RewardForwardFilter``).  There is no CPU fallback: importing works anywhere, computing needs a B200.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
