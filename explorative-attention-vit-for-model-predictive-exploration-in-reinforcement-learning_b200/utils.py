"""Drop-in for the numerics half of the reference's ``utils.py`` (make_train_data, RunningMeanStd,
RewardForwardFilter, global_grad_norm_, set_seed, Env_action_space_type, Logger) backed by the sm_100a
kernels.  numpy in / numpy out like the reference; device-resident variants avoid the host round trip."""
from __future__ import annotations

import contextlib
import logging
import os
import pickle
import random
import sys
import types
from enum import Enum
from typing import Optional

import numpy as np
import torch

from . import dist, ops
from .config import default_config


class Env_action_space_type(Enum):     # utils.py:34-36
    DISCRETE = 0
    CONTINUOUS = 1


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("eavit_b200 numerics run on CUDA (sm_100a) only; there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def make_train_data_device(reward: torch.Tensor, done: Optional[torch.Tensor], value: torch.Tensor, gamma: float,
                           lam: float):
    """GAE on device tensors.  reward f64 [E,T] + done u8 (extrinsic stream, train.py:748) or reward f32 and
    ``done=None`` (intrinsic, non-episodic stream, train.py:757).  Returns float64 (target, adv), flat [E*T]."""
    kind = 0 if reward.dtype == torch.float64 else 1
    return ops.gae_f64(reward.contiguous(), None if kind == 1 else done.contiguous(), value.contiguous(), gamma, lam, kind)


def make_train_data(reward, done, value, gamma, num_step, num_worker):
    """utils.py:42-67 (UseGAE branch), bit-identical float64 outputs incl. numpy's promotion quirks."""
    if not default_config.getboolean("UseGAE", fallback=True):
        raise NotImplementedError("UseGAE = False is not used by any reference config")
    lam = float(default_config["GAELambda"])
    reward, value = np.asarray(reward), np.asarray(value)
    assert reward.shape == (num_worker, num_step) and value.shape == (num_worker, num_step + 1)
    assert value.dtype == np.float32, "value buffers are float32 (train.py:593-594)"
    dev = _device()
    v = torch.from_numpy(np.ascontiguousarray(value)).to(dev)
    done = np.asarray(done)
    if reward.dtype == np.float64 and done.dtype == np.bool_:
        r = torch.from_numpy(np.ascontiguousarray(reward)).to(dev)
        d = torch.from_numpy(np.ascontiguousarray(done).view(np.uint8)).to(dev)
        ret, adv = ops.gae_f64(r, d, v, gamma, lam, 0)
    elif reward.dtype == np.float32 and done.dtype == np.float32:
        assert not done.any(), "the float32 `done` stream is np.zeros_like(reward) in the reference (train.py:759)"
        r = torch.from_numpy(np.ascontiguousarray(reward)).to(dev)
        ret, adv = ops.gae_f64(r, None, v, gamma, lam, 1)
    else:
        raise NotImplementedError(f"dtype combination reward={reward.dtype}, done={done.dtype} does not occur in train.py")
    return ret.cpu().numpy(), adv.cpu().numpy()


class RunningMeanStd(object):
    """utils.py:70-115.  Array-shaped statistics (obs_rms) live on the device as float64 and are updated by the
    kernels; scalar statistics (reward_rms) are three host doubles, exactly like the reference."""

    def __init__(self, epsilon=1e-4, shape=(), usage=""):
        assert usage in ["reward_rms", "obs_rms"], "Invalid usage param passed to RunningMeanStd"
        self.usage = usage
        self.shape = tuple(shape)
        self.train_method = default_config["TrainMethod"]
        if usage == "obs_rms":
            assert self.train_method in ["original_RND"], "hot path covers original_RND"
        self._on_device = len(self.shape) > 0
        if self._on_device:
            dev = _device()
            n = int(np.prod(self.shape))
            self._mean = torch.zeros(n, dtype=torch.float64, device=dev)
            self._var = torch.ones(n, dtype=torch.float64, device=dev)
            self._count = torch.full((1,), float(epsilon), dtype=torch.float64, device=dev)
        else:
            self._mean, self._var, self._count = np.zeros((), "float64"), np.ones((), "float64"), epsilon

    # public attributes of the reference (numpy views, D2H on access)
    @property
    def mean(self):
        return self._mean.cpu().numpy().reshape(self.shape) if self._on_device else self._mean

    @property
    def var(self):
        return self._var.cpu().numpy().reshape(self.shape) if self._on_device else self._var

    @property
    def count(self):
        return float(self._count.item()) if self._on_device else self._count

    def device_state(self):
        assert self._on_device
        return self._mean, self._var, self._count

    def update(self, x):
        """x: numpy or CUDA tensor [N, ...] (uint8 / float32 / float64)."""
        if not self._on_device:
            x = np.asarray(x)
            self.update_from_moments(np.mean(x, axis=0), np.var(x, axis=0), x.shape[0])
            return
        if isinstance(x, np.ndarray):
            x = torch.from_numpy(np.ascontiguousarray(x)).to(self._mean.device)
        x = x.contiguous()
        if dist.is_dist():
            # every rank holds its env shard: exchange (sum, sumsq about the shared current mean, count) -- SURVEY 8e(2)
            s, q = ops.rms_partial(x.view(x.shape[0], -1), self._mean)
            n = torch.tensor([float(x.shape[0])], dtype=torch.float64, device=s.device)
            dist.allreduce_sum_(s); dist.allreduce_sum_(q); dist.allreduce_sum_(n)
            ops.rms_merge(s, q, float(n.item()), self._mean, self._var, self._count)
            return
        ops.rms_update(x.view(x.shape[0], -1), self._mean, self._var, self._count)

    def update_from_moments(self, batch_mean, batch_var, batch_count):
        if self._on_device:
            m = torch.as_tensor(np.broadcast_to(np.asarray(batch_mean, dtype=np.float64), self.shape).reshape(-1).copy(),
                                device=self._mean.device)
            v = torch.as_tensor(np.broadcast_to(np.asarray(batch_var, dtype=np.float64), self.shape).reshape(-1).copy(),
                                device=self._mean.device)
            # sums about the current mean reproduce (mean, var) exactly in the Chan merge
            ds = m - self._mean
            ops.rms_merge(ds * batch_count, (v + ds * ds) * batch_count, float(batch_count), self._mean, self._var, self._count)
            return
        delta = batch_mean - self._mean
        tot_count = self._count + batch_count
        new_mean = self._mean + delta * batch_count / tot_count
        m2 = self._var * self._count + batch_var * batch_count + np.square(delta) * self._count * batch_count / tot_count
        self._mean, self._var, self._count = new_mean, m2 / tot_count, batch_count + self._count

    # checkpoints pickle these objects (train.py:939-941): store plain numpy under the reference's names
    def __getstate__(self):
        return dict(usage=self.usage, shape=self.shape, train_method=self.train_method, mean=np.array(self.mean),
                    var=np.array(self.var), count=self.count)

    def __setstate__(self, st):
        self.usage, self.train_method = st["usage"], st.get("train_method", "original_RND")
        mean, var = np.asarray(st["mean"], dtype=np.float64), np.asarray(st["var"], dtype=np.float64)
        self.shape = tuple(st.get("shape", mean.shape))
        self._on_device = len(self.shape) > 0
        if self._on_device:
            dev = _device()
            self._mean = torch.from_numpy(mean.reshape(-1).copy()).to(dev)
            self._var = torch.from_numpy(var.reshape(-1).copy()).to(dev)
            self._count = torch.full((1,), float(st["count"]), dtype=torch.float64, device=dev)
        else:
            self._mean, self._var, self._count = mean, var, st["count"]


def normalize_obs(x, obs_rms: RunningMeanStd, out_dtype=torch.float32) -> torch.Tensor:
    """train.py:666 / :855: ((x - mean) / sqrt(var)).clip(-5, 5) -> CUDA tensor (float32 like agents.py:212)."""
    mean, var, _ = obs_rms.device_state()
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x)).to(mean.device)
    x = x.contiguous()
    return ops.obs_normalize(x.view(x.shape[0], -1), mean, var, out_dtype=out_dtype).view(x.shape)


class RewardForwardFilter(object):
    """utils.py:118-128.  ``update`` keeps the per-step numpy API; ``filter_rollout`` runs a whole [E,T] rollout
    on the device (same float32 recurrence, bit-exact) and returns the moments train.py:738 needs."""

    def __init__(self, gamma):
        self.rewems = None
        self.gamma = gamma

    def update(self, rews):
        if self.rewems is None:
            self.rewems = rews
        else:
            self.rewems = self.rewems * self.gamma + rews
        return self.rewems

    def filter_rollout(self, int_reward: torch.Tensor) -> torch.Tensor:
        """int_reward CUDA float32 [E,T] -> moments float64 [5] (mean, var, T, sum, sumsq); updates ``rewems``."""
        E, T = int_reward.shape
        has = self.rewems is not None
        st = torch.as_tensor(np.asarray(self.rewems, dtype=np.float32), device=int_reward.device).contiguous() if has \
            else torch.zeros(E, dtype=torch.float32, device=int_reward.device)
        mom = ops.reward_filter(int_reward.contiguous(), st, has, float(self.gamma))
        self.rewems = st.cpu().numpy()
        if dist.is_dist():
            # global moments over all ranks' envs: all-reduce (sum, sumsq), n = T * E_total; count stays T (train.py:739)
            sums = mom[3:5].clone()
            dist.allreduce_sum_(sums)
            n = float(E * T * dist.world()[0])
            mean = sums[0] / n
            mom = torch.stack((mean, torch.clamp(sums[1] / n - mean * mean, min=0.0), mom[2], sums[0], sums[1]))
        return mom


def global_grad_norm_(parameters, norm_type=2):
    """utils.py:141-170 (L2 only): one fused reduction per flat buffer instead of 66 ``.item()`` syncs."""
    if isinstance(parameters, torch.Tensor):
        parameters = [parameters]
    grads = [p.grad for p in parameters if p.grad is not None]
    assert float(norm_type) == 2.0
    if not grads:
        return 0.0
    acc = torch.zeros(1, dtype=torch.float32, device=grads[0].device)
    for g in grads:
        n = g.numel()
        if g.is_contiguous() and n % 4 == 0 and g.data_ptr() % 16 == 0:
            ops.call("eavit_sumsq_f32", g, n, acc)
        else:
            acc += g.float().pow(2).sum()
    return float(acc.item()) ** 0.5


def set_seed(seed: int):               # utils.py:173-184
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)
    torch.backends.cudnn.deterministic = True


class Logger:
    """Minimal stand-in for utils.py:188-419 (observability is out of scope): console + optional file logging.
    ``RNDAgent`` asserts ``isinstance(logger, Logger)`` like the reference (agents.py:83)."""

    def __init__(self, file_log_path=None, tb_log_path=None):
        self.GLOBAL_RANK = int(os.environ.get("RANK", "0"))
        self.use_wandb = False
        self._log = logging.getLogger("eavit_b200")
        if file_log_path:
            os.makedirs(os.path.dirname(file_log_path + ".log") or ".", exist_ok=True)
            self._log.addHandler(logging.FileHandler(file_log_path + ".log", mode="w"))
        self.scalars = {}

    def log_msg_to_both_console_and_file(self, msg, only_rank_0=False):
        if not only_rank_0 or self.GLOBAL_RANK == 0:
            self._log.info(msg)

    def log_scalar_to_tb_without_step(self, tag, value, only_rank_0=False):
        self.scalars.setdefault(tag, []).append(float(value))

    def create_new_pytorch_profiler(self, *a, **k):
        pass

    def step_pytorch_profiler(self, *a, **k):
        pass


# ---------------------------------------------------------------------------------------------- checkpoints
# train.py:883-960 / :198-238: a checkpoint is one ``torch.save`` dict -- the agent's state_dicts plus the PICKLED
# ``obs_rms`` / ``reward_rms`` (RunningMeanStd) and ``discounted_reward`` (RewardForwardFilter) objects, which the pickle
# stream names by the reference's module path ``utils.<Class>``.  The two helpers keep that format loadable both ways:
# a reference checkpoint restores this package's device-backed objects, and a checkpoint written here restores the
# reference's own numpy classes (attribute for attribute: usage / mean / var / count / train_method, rewems / gamma).
_CKPT_CLASSES = ("RunningMeanStd", "RewardForwardFilter")


class _RefUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module == "utils" and name in _CKPT_CLASSES:
            return globals()[name]
        return super().find_class(module, name)


_ref_pickle = types.ModuleType("eavit_b200_ref_pickle")
_ref_pickle.__dict__.update({k: getattr(pickle, k) for k in dir(pickle) if not k.startswith("__")})
_ref_pickle.Unpickler = _RefUnpickler
_ref_pickle.load = lambda f, **kw: _RefUnpickler(f, **kw).load()


def load_checkpoint(path, map_location=None) -> dict:
    """``torch.load`` of a checkpoint written by the reference (train.py:926-959) or by ``save_checkpoint``: tensors as
    usual, ``utils.RunningMeanStd`` / ``utils.RewardForwardFilter`` pickles come back as this package's classes."""
    return torch.load(path, map_location=map_location, weights_only=False, pickle_module=_ref_pickle)


@contextlib.contextmanager
def _reference_class_paths():
    """Make ``utils.RunningMeanStd`` / ``utils.RewardForwardFilter`` resolvable while pickling, without importing the
    reference: plain attribute-bag classes under that module path (pickle stores the path + ``__dict__`` only)."""
    prev = sys.modules.get("utils")
    mod = prev
    added = []
    if mod is None:
        mod = types.ModuleType("utils")
        sys.modules["utils"] = mod
    for name in _CKPT_CLASSES:
        if not hasattr(mod, name):
            setattr(mod, name, type(name, (object,), {"__module__": "utils", "__qualname__": name}))
            added.append(name)
    try:
        yield mod
    finally:
        for name in added:
            delattr(mod, name)
        if prev is None:
            del sys.modules["utils"]


def _as_reference_object(obj, mod):
    if isinstance(obj, RunningMeanStd):
        cls = getattr(mod, "RunningMeanStd")
        if isinstance(obj, cls):                       # this package already IS the importable ``utils``
            return obj
        st = obj.__getstate__()
        st.pop("shape")
        ref = cls.__new__(cls)
        ref.__dict__.update(st)
        return ref
    if isinstance(obj, RewardForwardFilter):
        cls = getattr(mod, "RewardForwardFilter")
        if isinstance(obj, cls):
            return obj
        ref = cls.__new__(cls)
        ref.__dict__.update(rewems=None if obj.rewems is None else np.array(obj.rewems), gamma=obj.gamma)
        return ref
    return obj


def save_checkpoint(ckpt_dict: dict, path) -> None:
    """``torch.save(ckpt_dict, path)`` (train.py:958-959) with the statistics objects written in the reference's own
    pickle format, so that ``torch.load`` inside the reference's train.py:198-238 restores its numpy classes."""
    with _reference_class_paths() as mod:
        torch.save({k: _as_reference_object(v, mod) for k, v in ckpt_dict.items()}, path)
