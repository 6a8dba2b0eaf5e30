"""Device-resident rollout buffer: the train.py:582-599 lists, the per-step appends of train.py:676-687
and the post-processing of train.py:707-779 / :855, kept on the GPU.

The reference appends float64 numpy arrays step-major, transposes them to env-major on the host (3.7 GB of copies at 128
envs) and re-materialises everything as torch tensors per minibatch.  Here every step is written straight into its
env-major slot ``e * T + t`` (the sample index of agents.py:301), frames stay uint8 (they are raw ALE frames: exact), and
``finish()`` runs the existing kernels -- reward filter + moments, reward normalisation, both GAE streams (float64,
numpy-promotion exact), advantage combine, observation statistics update and normalisation -- and returns the
``RNDAgent.train_model`` argument tuple as CUDA tensors (no host round trip; train_model accepts them as they are).
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import ops
from .utils import RewardForwardFilter, RunningMeanStd, _device, make_train_data_device, normalize_obs


def _dev(x, device, dtype=None) -> torch.Tensor:
    t = x if torch.is_tensor(x) else torch.from_numpy(np.ascontiguousarray(x))
    t = t.to(device, non_blocking=True)
    return t if dtype is None or t.dtype == dtype else t.to(dtype)


class DeviceRollout:
    """One update's worth of experience for ``num_env`` environments x ``num_step`` steps.

    ``add(t, ...)`` takes what train.py:676-687 appends at step ``t`` (numpy or CUDA tensors, reference dtypes):
    states / next_obs are frames with values 0..255 (uint8, or the reference's float arrays holding the same integers),
    reward float64 (unclipped), done bool, action int64, value_ext / value_int float32, policy = raw logits float32
    [E, A], int_reward float32.  ``add_last_values`` is the extra get_action of train.py:702-704."""

    def __init__(self, num_env: int, num_step: int, n_actions: int, image: int = 84, stack: int = 4, device=None):
        self.E, self.T, self.A = num_env, num_step, n_actions
        dev = self.device = torch.device(device) if device is not None else _device()
        E, T = num_env, num_step
        self.states = torch.empty(E, T, stack, image, image, dtype=torch.uint8, device=dev)
        self.next_obs = torch.empty(E, T, 1, image, image, dtype=torch.uint8, device=dev)
        self.reward = torch.empty(E, T, dtype=torch.float64, device=dev)
        self.done = torch.empty(E, T, dtype=torch.uint8, device=dev)
        self.action = torch.empty(E, T, dtype=torch.int64, device=dev)
        self.value_ext = torch.empty(E, T + 1, dtype=torch.float32, device=dev)
        self.value_int = torch.empty(E, T + 1, dtype=torch.float32, device=dev)
        self.policy = torch.empty(E, T, n_actions, dtype=torch.float32, device=dev)
        self.int_reward = torch.empty(E, T, dtype=torch.float32, device=dev)

    @staticmethod
    def _frames(x, device) -> torch.Tensor:
        t = _dev(x, device)
        if t.dtype != torch.uint8:                      # float frames of the reference: integers 0..255, exact in uint8
            t = t.to(torch.uint8)
        return t

    def add(self, t: int, states, next_obs, reward, done, action, value_ext, value_int, policy, int_reward) -> None:
        assert 0 <= t < self.T
        d = self.device
        self.states[:, t] = self._frames(states, d)
        self.next_obs[:, t] = self._frames(next_obs, d).view(self.E, 1, *self.next_obs.shape[-2:])
        self.reward[:, t] = _dev(reward, d, torch.float64)
        self.done[:, t] = _dev(done, d).to(torch.uint8)
        self.action[:, t] = _dev(action, d, torch.int64)
        self.value_ext[:, t] = _dev(value_ext, d, torch.float32)
        self.value_int[:, t] = _dev(value_int, d, torch.float32)
        self.policy[:, t] = _dev(policy, d, torch.float32)
        self.int_reward[:, t] = _dev(int_reward, d, torch.float32)

    def add_last_values(self, value_ext, value_int) -> None:
        self.value_ext[:, self.T] = _dev(value_ext, self.device, torch.float32)
        self.value_int[:, self.T] = _dev(value_int, self.device, torch.float32)

    def finish(self, obs_rms: RunningMeanStd, reward_rms: RunningMeanStd, reward_filter: RewardForwardFilter, gamma: float,
               int_gamma: float, lam: float, ext_coef: float, int_coef: float):
        """train.py:707-779 + :855.  Returns ``(states u8 [N,C,H,W], target_ext f64 [N], target_int f64 [N], action i64 [N],
        adv f64 [N], next_obs_norm f32 [N,1,H,W], old_policy f32 [N,A])`` -- all CUDA, env-major (index e*T + t) -- in the
        argument order of ``RNDAgent.train_model``.  Updates ``reward_filter``, ``reward_rms`` and ``obs_rms`` in place,
        in the reference's order (the rollout itself used the statistics from BEFORE this update, train.py:666)."""
        E, T = self.E, self.T
        N = E * T
        # intrinsic reward normalisation (train.py:736-743): forward filter per env, moments over the [T,E] outputs, count = T
        mom = reward_filter.filter_rollout(self.int_reward).cpu().numpy()
        reward_rms.update_from_moments(float(mom[0]), float(mom[1]), int(round(float(mom[2]))))
        int_reward = self.int_reward.clone()
        ops.scale_by_rsqrt_var(int_reward, torch.tensor([float(reward_rms.var)], dtype=torch.float64, device=self.device))
        # both GAE streams (train.py:748-760): extrinsic is episodic (clipped float64 reward + done), intrinsic is not
        ext_target, ext_adv = make_train_data_device(self.reward.clamp(-1.0, 1.0), self.done, self.value_ext, gamma, lam)
        int_target, int_adv = make_train_data_device(int_reward, None, self.value_int, int_gamma, lam)
        adv = ops.axpby_f64(int_adv, ext_adv, float(int_coef), float(ext_coef))               # train.py:767
        next_obs = self.next_obs.view(N, 1, *self.next_obs.shape[-2:])
        obs_rms.update(next_obs)                                                               # train.py:774
        obs_norm = normalize_obs(next_obs, obs_rms)                                            # train.py:855 (after the update)
        return (self.states.view(N, *self.states.shape[2:]), ext_target, int_target, self.action.view(N), adv, obs_norm,
                self.policy.view(N, self.A))
