"""Drop-in for the reference's ``agents.py``: ``RNDAgent`` with the same constructor, attributes and the three
hot calls ``get_action`` / ``compute_intrinsic_reward`` / ``train_model`` (agents.py:30-535), running on the
sm_100a kernels.

What changes underneath (results stay within the north-star tolerances):
  * the rollout is uploaded ONCE per update and stays resident on the device; minibatches are gathered by
    index inside the kernels (the reference re-materialises the whole rollout per minibatch, agents.py:288-301);
  * all trainable tensors live in one flat buffer: one fused Adam launch, one NCCL all-reduce per step;
  * no per-step host syncs: loss terms are accumulated on the device and read back once per update;
  * data-parallel: with torch.distributed initialised, every rank trains on its env shard and the flat gradient
    is all-reduced (mean) before Adam -- the semantics the reference's DDP wrapper intended.
"""
from __future__ import annotations

import weakref
from typing import Optional

import os

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, dist, ops
from .config import default_config
from .model import CnnActorCriticNetwork, RNDModel, Runtime, ViT_IMPLEMENTATION, _as_device_image
from .ops import call
from .utils import Env_action_space_type, Logger


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam(lr) semantics (agents.py:129) as one fused launch over the agent's flat parameter store.

    ``state_dict`` / ``load_state_dict`` speak torch.optim.Adam's own layout (``state[i] = {step, exp_avg, exp_avg_sq}``,
    ``param_groups[0]['params'] = [0..n)``), so the checkpoint entry ``agent.optimizer.state_dict`` (train.py:931) is a
    dict ``torch.optim.Adam.load_state_dict`` accepts and vice versa.  Index i is the i-th tensor of the flat store
    (``eavit_param_names`` records the names).  The reference builds its optimiser from a *set* of parameters
    (agents.py:141-164), so the index -> tensor mapping of a reference-written state is the hash order of one process and
    is recorded nowhere: such a state is matched by shape, in store order among equal shapes, with a warning."""

    _TORCH_GROUP_DEFAULTS = dict(amsgrad=False, maximize=False, foreach=None, capturable=False, differentiable=False, fused=None)

    def __init__(self, params, lr, agent):
        super().__init__(list(params), dict(lr=lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0))
        self._agent = weakref.ref(agent)

    def zero_grad(self, set_to_none: bool = False):
        self._agent().runtime().store.zero_grad()

    @torch.no_grad()
    def step(self, closure=None):
        g = self.param_groups[0]
        ag = self._agent()
        rt = ag.runtime()
        rt.store.adam_step(g["lr"], 1.0, g["betas"][0], g["betas"][1], g["eps"], ranges=ag._trainable_ranges(rt))
        rt.refresh_after_step()

    def state_dict(self):
        st = self._agent().runtime().store
        names = list(st.shapes.keys())
        step = st.step.to(torch.float32).cpu().reshape(())                 # torch keeps `step` as a float32 CPU scalar tensor
        state = {i: {"step": step.clone(), "exp_avg": st._view(st.m, n).clone(), "exp_avg_sq": st._view(st.v, n).clone()}
                 for i, n in enumerate(names)}
        group = {k: v for k, v in self.param_groups[0].items() if k != "params"}
        for k, v in self._TORCH_GROUP_DEFAULTS.items():
            group.setdefault(k, v)
        group["params"] = list(range(len(names)))
        return {"state": state, "param_groups": [group], "eavit_param_names": names}

    def load_state_dict(self, sd):
        st = self._agent().runtime().store
        names = list(st.shapes.keys())
        if "flat_exp_avg" in sd:                                           # round-1 private format
            if list(sd.get("names", names)) != names or sd["flat_exp_avg"].numel() != st.numel:
                raise ValueError("FusedAdam.load_state_dict: flat optimiser state was written for a different tensor layout")
            st.m.copy_(sd["flat_exp_avg"]); st.v.copy_(sd["flat_exp_avg_sq"]); st.step.copy_(sd["step"])
            return
        state, groups = sd["state"], sd["param_groups"]
        saved = sd.get("eavit_param_names")
        if saved is not None:
            if len(saved) != len(groups[0]["params"]):
                raise ValueError("FusedAdam.load_state_dict: eavit_param_names does not match param_groups")
            index_of = {n: i for n, i in zip(saved, groups[0]["params"])}
            unknown = [n for n in saved if n not in st.shapes]
            if unknown:
                raise ValueError(f"FusedAdam.load_state_dict: state names unknown to this agent: {unknown[:4]}")
        else:
            # reference-written state: indices follow that process's set order; match by shape, store order within a shape
            by_shape = {}
            for i in groups[0]["params"]:
                e = state.get(i)
                if e is not None:
                    by_shape.setdefault(tuple(e["exp_avg"].shape), []).append(i)
            index_of = {}
            ambiguous = False
            for n in names:
                cand = by_shape.get(tuple(st.shapes[n]), [])
                if cand:
                    ambiguous |= len(cand) > 1
                    index_of[n] = cand.pop(0)
            if ambiguous:
                self._agent().logger.log_msg_to_both_console_and_file(
                    "FusedAdam.load_state_dict: optimiser state without parameter names (written by the reference from a set of "
                    "parameters): Adam moments of equal-shaped tensors were assigned in store order", only_rank_0=True)
        step = 0.0
        loaded = 0
        for n in names:
            e = state.get(index_of.get(n, -1))
            if e is None:
                continue                                                   # a tensor the writer never stepped (frozen)
            if tuple(e["exp_avg"].shape) != tuple(st.shapes[n]):
                raise ValueError(f"FusedAdam.load_state_dict: {n}: exp_avg shape {tuple(e['exp_avg'].shape)} != {tuple(st.shapes[n])}")
            st._view(st.m, n).copy_(e["exp_avg"])
            st._view(st.v, n).copy_(e["exp_avg_sq"])
            step = max(step, float(e["step"]))
            loaded += 1
        if state and not loaded:
            raise ValueError("FusedAdam.load_state_dict: no state entry matches a tensor of this agent")
        st.step.fill_(int(step))
        g = self.param_groups[0]
        for k in ("lr", "betas", "eps", "weight_decay"):
            if k in groups[0]:
                g[k] = tuple(groups[0][k]) if k == "betas" else groups[0][k]


class _CapturedCall:
    """fn(static device input) -> device tensor, captured once as a CUDA graph (fixed shapes, buffers and weights)."""

    def __init__(self, fn, example: torch.Tensor):
        self.x = example.clone()
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream(device=example.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):                    # warm-up: scratch buffers, smem attributes, epoch word
            for _ in range(2):
                fn(self.x)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count(include_replays=False)
        with torch.cuda.graph(self.graph):
            self.out = fn(self.x)
        self.nodes = _lib.launch_count(include_replays=False) - n0

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        self.x.copy_(x, non_blocking=True)
        self.graph.replay()
        _lib.add_replayed(self.nodes)
        return self.out


class RNDAgent(nn.Module):
    def __init__(self, input_size, output_size, env_action_space_type, num_env, num_step, gamma, GAE_Lambda=0.95,
                 learning_rate=1e-4, ent_coef=0.01, max_grad_norm=0.5, epoch=3, batch_size=128, ppo_eps=0.1,
                 update_proportion=0.25, use_gae=True, use_cuda=False, use_noisy_net=False,
                 representation_lr_method="BYOL", device=None, logger: Logger = None):
        super().__init__()
        self.env_action_space_type = env_action_space_type
        vt = int(default_config["ViT_implementation_type"])
        ViT_implementation_type = ViT_IMPLEMENTATION(vt)      # 0 lucidrains, 1 HF-style, 2 = the original RND CNN backbone
        self.model = CnnActorCriticNetwork(input_size, output_size, env_action_space_type, use_noisy_net,
                                           ViT_implementation_type=ViT_implementation_type)
        self.num_env, self.output_size, self.input_size, self.num_step = num_env, output_size, input_size, num_step
        self.gamma, self.GAE_Lambda, self.epoch, self.batch_size = gamma, GAE_Lambda, epoch, batch_size
        self.use_gae, self.ent_coef, self.ppo_eps, self.max_grad_norm = use_gae, ent_coef, ppo_eps, max_grad_norm
        self.update_proportion = update_proportion
        self.use_cuda = use_cuda
        if not use_cuda:
            raise RuntimeError("eavit_b200.RNDAgent requires use_cuda=True and a CUDA device: the learner hot path is "
                               "hand-written sm_100a CUDA with no CPU fallback")
        self.device = torch.device(device if device is not None else "cuda")
        assert isinstance(logger, Logger)
        self.logger = logger
        self.train_method = default_config["TrainMethod"]
        assert self.train_method == "original_RND", "hot path covers TrainMethod = original_RND (SURVEY fact 7)"
        self.rnd = RNDModel(input_size=input_size, output_size=512, train_method=self.train_method)
        assert representation_lr_method == "None", "SSL heads are out of scope (incompatible with the ViT backbone, SURVEY fact 7)"
        self.representation_lr_method = representation_lr_method
        self.representation_model = None
        self.representation_loss_coef = 0
        self.model = self.model.to(self.device)
        self.rnd = self.rnd.to(self.device)
        self._rt: Optional[Runtime] = None
        ref = weakref.ref(self)
        self.model._agent_rt = lambda: ref().runtime()
        self.rnd._agent_rt = lambda: ref().runtime()
        self.optimizer = FusedAdam(self.get_agent_parameters(), learning_rate, self)
        self.freeze_shared_backbone_during_training = default_config.getboolean("freeze_shared_backbone", fallback=False)
        self.world_size, self.rank = 1, 0
        self.last_stats = None
        self._ws = {}

    # ---- reference API ---------------------------------------------------------------------------------
    def get_agent_parameters(self):
        """agents.py:141-164: PPO model + RND predictor parameters, as a set."""
        return set([*self.model.parameters(), *self.rnd.predictor.parameters()])

    def set_mode(self, mode="train"):
        assert mode in ["train", "eval"]
        self.model = self.model.train(mode == "train")
        self.rnd = self.rnd.train(mode == "train")

    def runtime(self) -> Runtime:
        if self._rt is None or not self._rt.valid():
            old = self._rt
            self._rt = Runtime(self, "", n_actions=self.output_size,
                               ext_uses_int_critic=self.model.ViT_implementation_type == ViT_IMPLEMENTATION.HG_ViT)
            if old is not None and not self._rt.store.adopt_optimizer_state(old.store):
                # a rebuild (module.to(), load_state_dict(assign=True)) must not silently restart Adam from zero moments
                raise RuntimeError("eavit_b200.RNDAgent: the parameter layout changed under a live optimiser; "
                                   "rebuild the agent (or reload the optimiser state) instead")
            self._frozen_key = None
            if dist.is_dist():
                self.world_size, self.rank = dist.world()
                # start every rank from rank 0's weights (what DDP's constructor did at train.py:243)
                dist.broadcast_(self._rt.store.flat, 0)
                dist.broadcast_(self._rt.frozen.flat, 0)
                self._rt.sync()
        return self._rt

    # ---- rollout side: one captured CUDA graph per call shape -------------------------------------
    # get_action / compute_intrinsic_reward run once per env step on E samples: ~45 / ~25 small launches whose host cost
    # exceeds their GPU time.  Each is captured once per (batch, dtype) and replayed: H2D into the static input, one
    # graph launch, one packed D2H.  With dropout active (the reference rolls out in train mode, fact 6) the captured
    # seeds are constants, so the graph starts with a dropout-epoch bump and every replay draws fresh masks.
    # EAVIT_ROLLOUT_GRAPH=0 (or an active kernel profile) keeps the eager launches.
    def _graphed(self, kind: str, x: torch.Tensor, fn):
        rt = self.runtime()
        if os.environ.get("EAVIT_ROLLOUT_GRAPH", "1") != "1" or ops._PROF is not None:
            return fn(x)
        if getattr(self, "_graph_rt", None) is not rt:
            self._graphs, self._graph_rt = {}, rt
        key = (kind, tuple(x.shape), x.dtype, bool(self.model.training))
        g = self._graphs.get(key)
        if g is None:
            g = self._graphs[key] = _CapturedCall(fn, x)
        # the replay overwrites the (kind, batch) scratch activations and, with dropout, bumps the device epoch word:
        # a pending autograd backward of the same buffers must notice (model.Runtime.check_gen)
        rt._bump_gen("ac" if kind == "act" else "rnd", x.shape[0])
        if kind == "act" and rt.dropout_active():
            rt._epoch_gen = getattr(rt, "_epoch_gen", 0) + 1
        return g(x)

    def _act_device(self, x: torch.Tensor) -> torch.Tensor:
        rt = self.runtime()
        if rt.dropout_active():
            call("eavit_dropout_epoch_bump")
            rt._epoch_gen = getattr(rt, "_epoch_gen", 0) + 1
        pol, ve, vi = rt.ac_forward(x, x.shape[0])
        return torch.cat((pol.reshape(-1), ve, vi))                       # one packed D2H read instead of four

    def _rnd_device(self, x: torch.Tensor) -> torch.Tensor:
        rt = self.runtime()
        x = x.to(torch.float32)                                           # torch.FloatTensor(next_obs), agents.py:212
        B = x.shape[0]
        tgt = rt.rnd_tgt.forward(x, B)
        prd = rt.rnd_pred.forward(x, B, col0=rt.rnd_tgt.buf[B].t["col0"])
        return ops.intrinsic_mse(tgt, prd)

    @torch.no_grad()
    def get_action(self, state):
        """agents.py:187-195 (DISCRETE): state float32 [E,C,H,W] (already /255) or uint8 raw frames.
        Returns (action int64 [E], value_ext f32 [E], value_int f32 [E], logits f32 [E,A]) as numpy."""
        rt = self.runtime()
        rt.sync_if_changed()
        x = _as_device_image(state, rt.device)
        E, A = x.shape[0], self.output_size
        pack = self._graphed("act", x, self._act_device).cpu().numpy()
        policy = pack[: E * A].reshape(E, A).copy()
        value_ext, value_int = pack[E * A: E * A + E].copy(), pack[E * A + E:].copy()
        z = policy - policy.max(axis=1, keepdims=True)
        e = np.exp(z, dtype=np.float32)
        action_prob = e / e.sum(axis=1, keepdims=True)                    # F.softmax(policy, dim=-1) in float32
        action = self.random_choice_prob_index(action_prob)
        return action, value_ext.squeeze(), value_int.squeeze(), policy

    @staticmethod
    def random_choice_prob_index(p, axis=1):                              # agents.py:205-208, host numpy RNG
        r = np.expand_dims(np.random.rand(p.shape[1 - axis]), axis=axis)
        return (p.cumsum(axis=axis) > r).argmax(axis=axis)

    @torch.no_grad()
    def compute_intrinsic_reward(self, next_obs):
        """agents.py:210-218: next_obs float64/float32 numpy (or CUDA tensor) [E,1,H,W], already normalised."""
        rt = self.runtime()
        rt.sync_if_changed()
        x = next_obs if torch.is_tensor(next_obs) else torch.from_numpy(np.ascontiguousarray(next_obs))
        x = x.to(rt.device).contiguous()
        return self._graphed("rnd", x, self._rnd_device).cpu().numpy()

    # ---- update ----------------------------------------------------------------------------------------
    @staticmethod
    def _rnd_grad_range(st):
        """[lo, hi) of the RND predictor's tensors in the flat store, or None when they are not one contiguous block."""
        cached = getattr(st, "_rnd_range", False)
        if cached is not False:
            return cached
        offs = sorted((o, n) for n, o in st.offsets.items())
        ends = [o for o, _ in offs[1:]] + [st.numel]
        rnd = [(o, e) for (o, n), e in zip(offs, ends) if n.startswith("rnd.predictor.")]
        rng = None
        if rnd:
            lo, hi = rnd[0][0], rnd[-1][1]
            inside = [n for (o, n) in offs if lo <= o < hi]
            if all(n.startswith("rnd.predictor.") for n in inside) and len(inside) == len(rnd):
                rng = (lo, hi)
        st._rnd_range = rng
        return rng

    def _trainable_ranges(self, rt):
        """None when every tensor of the store trains; else the merged flat ranges of the tensors that do.  Frozen =
        ``requires_grad is False`` on the Parameter -- what train.py:261-263 sets on ``model.feature.*`` when the config
        says ``freeze_shared_backbone = True`` (torch.optim.Adam then skips those tensors, nn.utils.clip_grad_norm_ and
        global_grad_norm_ ignore them)."""
        key = tuple(p.requires_grad for n, p in rt.params.items() if n in rt.store.shapes)
        if getattr(self, "_frozen_key", None) != key:
            frozen = rt.frozen_names()
            self._frozen_key = key
            self._frozen_ranges = rt.store.name_ranges(frozen) if frozen else None
            self._train_ranges = rt.store.name_ranges([n for n in rt.store.shapes if n not in set(frozen)]) if frozen else None
            self._backbone_frozen = bool(frozen) and all(
                (not p.requires_grad) for n, p in rt.params.items() if n.startswith("model.feature."))
        return self._train_ranges

    def _layer_ranges(self, rt, li):
        cache = getattr(rt, "_layer_range_cache", None)
        if cache is None:
            cache = rt._layer_range_cache = {}
        if li not in cache:
            cache[li] = rt.store.name_ranges(rt.encoder.layer_param_names(li))
        return cache[li]

    @staticmethod
    def _complement(ranges, n):
        """[0, n) minus the given disjoint [lo, hi) ranges, as merged ranges."""
        out, pos = [], 0
        for lo, hi in sorted(ranges):
            if lo > pos:
                out.append((pos, lo))
            pos = max(pos, hi)
        if pos < n:
            out.append((pos, n))
        return out

    def _side_stream(self, rt):
        s = getattr(rt, "_side", None)
        if s is None:
            s = rt._side = torch.cuda.Stream(device=rt.device)
        return s

    def _scratch(self, B, A, dev):
        key = (B, A)
        w = self._ws.get(key)
        if w is None:
            f = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
            w = dict(te=f(B), ti=f(B), adv=f(B), y=torch.empty(B, dtype=torch.int64, device=dev), old=f(B, A),
                     dpol=f(B, A), dv=f(2 * B),
                     dpred=torch.empty(B, 512, dtype=torch.bfloat16, device=dev), idx=torch.empty(B, dtype=torch.int64, device=dev),
                     mask=f(B))
            both = torch.zeros(32, dtype=torch.float32, device=dev)
            w["stats"], w["rnd_stats"] = both[:16], both[16:]
            self._ws[key] = w
        return w

    def upload_rollout(self, states, target_ext, target_int, y, adv, next_obs_norm, old_policy):
        """Host numpy (reference dtypes) -> device-resident tensors.  CUDA tensors pass through untouched."""
        dev = self.runtime().device

        def up(a, dt=None):
            t = a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a))
            t = t.to(dev, non_blocking=True)
            return t if dt is None or t.dtype == dt else t.to(dt)
        st = up(states)
        if st.dtype not in (torch.uint8, torch.float32):
            st = st.float()
        A = old_policy.shape[-1]
        old = up(old_policy, torch.float32)
        if old.dim() == 3:                                               # [T,E,A] -> [E*T, A]  (agents.py:301)
            old = old.permute(1, 0, 2).contiguous().view(-1, A)
        fresh = dict(states=st.contiguous(), te=up(target_ext, torch.float32), ti=up(target_int, torch.float32),
                     y=up(y, torch.int64), adv=up(adv, torch.float32), obs=up(next_obs_norm, torch.float32).contiguous(),
                     old=old.contiguous())
        # The rollout lives in buffers that keep their addresses from update to update (same shapes / dtypes): the captured
        # step graph stays valid across updates.  One device-to-device copy of what was uploaded (0.1 ms at cfg3).
        pool = self.__dict__.setdefault("_rollout_buf", {})              # one buffer per (name, shape, dtype): raw-frame and float
        theirs = {a.data_ptr() for a in (states, target_ext, target_int, y, adv, next_obs_norm, old_policy) if torch.is_tensor(a)}
        R = {}                                                           # callers may alternate without invalidating the graphs
        for k, t in fresh.items():
            sig = (k, tuple(t.shape), t.dtype, t.device)
            b = pool.get(sig)
            if b is None:
                if len(pool) >= 21:                                      # shapes changed for good: let the old buffers go
                    pool.clear()
                b = pool[sig] = t.clone() if t.data_ptr() in theirs else t   # never adopt (and later overwrite) the caller's tensor
            else:
                b.copy_(t, non_blocking=True)
            R[k] = b
        return R

    # ---- the optimiser step as ONE captured CUDA graph ---------------------------------------------------------------
    # A step is ~150 kernel launches on two streams.  Their arguments are fixed for a given (rollout buffers, minibatch size,
    # optimiser constants): the minibatch indices, the RND mask and the Adam step counter are DEVICE data.  So the step is
    # captured once and replayed -- the host cost of a step drops from ~1.3 ms of Python / ctypes / tensor-map encoding to
    # two small copies and one graph launch, and the kernels run back to back without launch gaps.  Dropout: the captured
    # seeds are constants, so the graph starts by bumping the device epoch word that every mask folds in (fresh masks per
    # replay, the same word for the forward and the backward of one replay).  EAVIT_STEP_GRAPH=0 keeps eager launches;
    # ``apply=False`` (gradient inspection), an active kernel profile and data-parallel runs run eagerly as well: capturing
    # the NCCL exchange (EAVIT_STEP_GRAPH_DIST=1) hung at 2 ranks with the NCCL 2.28 of this image, so it stays opt-in.
    def train_step(self, R: dict, idx: torch.Tensor, mask: torch.Tensor, stats_out: Optional[torch.Tensor] = None,
                   apply: bool = True):
        """One minibatch: agents.py:284-508 from the batch gather to ``optimizer.step()``."""
        if (apply and ops._PROF is None and os.environ.get("EAVIT_STEP_GRAPH", "1") == "1"
                and (self.world_size == 1 or os.environ.get("EAVIT_STEP_GRAPH_DIST", "0") == "1")
                and not torch.cuda.is_current_stream_capturing()):
            return self._train_step_graphed(R, idx, mask, stats_out)
        return self._train_step_eager(R, idx, mask, stats_out, apply)

    def _train_step_graphed(self, R, idx, mask, stats_out):
        rt = self.runtime()
        B = idx.numel()
        g = self.optimizer.param_groups[0]
        key = (id(rt), B, tuple((k, v.data_ptr(), v.dtype) for k, v in sorted(R.items())), bool(self.model.training),
               rt.dropout_active(), tuple(p.requires_grad for n, p in rt.params.items() if n in rt.store.shapes),
               float(g["lr"]), tuple(g["betas"]), float(g["eps"]), default_config.getboolean("UseGradClipping", fallback=False),
               float(self.max_grad_norm), float(self.ppo_eps), float(self.ent_coef), self.world_size)
        cache = self.__dict__.setdefault("_step_graphs", {})
        ent = cache.get(key)
        if ent is None:
            if len(cache) >= 4:                                         # a new rollout allocation: drop the stale captures
                cache.clear()
            dev = rt.device
            sidx = torch.empty(B, dtype=torch.int64, device=dev)
            smask = torch.empty(B, dtype=torch.float32, device=dev)
            sstats = torch.zeros(16, dtype=torch.float32, device=dev)
            sidx.copy_(idx); smask.copy_(mask)
            # one eager pass WITHOUT the optimiser update: allocates every scratch buffer, uploads the geometry tables and sets the
            # kernel attributes (none of which may happen inside a capture); its gradients are discarded by the next zero_grad
            self._train_step_eager(R, sidx, smask, sstats, False)
            torch.cuda.synchronize(dev)
            # capture_begin / capture_end directly: the torch.cuda.graph context manager also runs gc.collect() and
            # empty_cache(), which costs ~0.5 s next to a 3 GB rollout and makes the next upload re-cudaMalloc its buffers
            graph = torch.cuda.CUDAGraph()
            cur = torch.cuda.current_stream(dev)
            cap = self.__dict__.get("_capture_stream") or torch.cuda.Stream(device=dev)
            self._capture_stream = cap
            cap.wait_stream(cur)
            n0 = _lib.launch_count(include_replays=False)
            with torch.cuda.stream(cap):
                graph.capture_begin()
                try:
                    if rt.dropout_active():
                        call("eavit_dropout_epoch_bump")
                    self._train_step_eager(R, sidx, smask, sstats, True)
                finally:
                    graph.capture_end()
            cur.wait_stream(cap)
            nodes = _lib.launch_count(include_replays=False) - n0      # kernel nodes of the graph = launches of one replay
            ent = cache[key] = (graph, sidx, smask, sstats, nodes)
        graph, sidx, smask, sstats, nodes = ent
        sidx.copy_(idx, non_blocking=True)
        smask.copy_(mask, non_blocking=True)
        graph.replay()
        _lib.add_replayed(nodes)
        rt._bump_gen("ac", B)                                           # a pending autograd backward of these buffers must notice
        if rt.dropout_active():
            rt._epoch_gen = getattr(rt, "_epoch_gen", 0) + 1
        if stats_out is not None:
            stats_out.copy_(sstats, non_blocking=True)

    def _train_step_eager(self, R: dict, idx: torch.Tensor, mask: torch.Tensor, stats_out: Optional[torch.Tensor] = None,
                          apply: bool = True):
        rt = self.runtime()
        B, A = idx.numel(), self.output_size
        w = self._scratch(B, A, rt.device)
        st = rt.store
        gs = 1.0
        st.zero_grad()
        call("eavit_zero", w["stats"], 128)                                 # stats | rnd_stats (adjacent)
        call("eavit_gather_batch", idx, B, A, R["te"], R["ti"], R["adv"], R["y"], R["old"], w["te"], w["ti"], w["adv"], w["y"], w["old"])
        # RND (agents.py:333-338): target is frozen, predictor gets the masked MSE gradient
        # The RND towers (~50 small launches, < 1 % of the FLOPs) and the ViT + heads are independent until the optimiser
        # step: the towers run on a second stream so that their launch-latency-bound kernels fill the tails of the big
        # ViT kernels instead of sitting serially in front of them.  Their loss term goes to its own stats slot (5) and
        # their gradients to the predictor's slice of the flat gradient, so the two streams never write the same bytes.
        early = None
        cur = torch.cuda.current_stream()
        side = self._side_stream(rt) if ops._PROF is None and os.environ.get("EAVIT_RND_STREAM", "1") == "1" else None
        if side is not None:
            side.wait_stream(cur)
            torch.cuda.set_stream(side)
        try:
            pred = rt.rnd_pred.forward(R["obs"], B, idx)
            tgt = rt.rnd_tgt.forward(R["obs"], B, idx, col0=rt.rnd_pred.buf[B].t["col0"])
            call("eavit_rnd_loss", pred, tgt, mask, B, pred.shape[1], gs, w["dpred"], None, w["rnd_stats"])
            rt.rnd_pred.backward(w["dpred"])
            early = self._rnd_grad_range(st) if (side is not None and self.world_size > 1) else None
            if early is not None:
                # the predictor's gradient is complete long before the ViT backward ends: exchange its slice of the flat
                # buffer now, on the towers' stream, so that only the ViT + heads slices are left for the end of the step
                dist.allreduce_sum_(st.grad[early[0]:early[1]])
        finally:
            if side is not None:
                torch.cuda.set_stream(cur)
        # PPO (agents.py:455-494)
        pol, ve, vi = rt.ac_forward(R["states"], B, idx)
        call("eavit_ppo_loss", pol, w["old"], w["y"], w["adv"], ve, vi, w["te"], w["ti"], B, A, float(self.ppo_eps),
             float(self.ent_coef), gs, w["dpol"], w["dv"][B:], w["dv"][:B], w["stats"])
        train_ranges = self._trainable_ranges(rt)
        # Data-parallel (replaces the reference's never-armed DDP reducer, train.py:243).  Optional (EAVIT_GRAD_OVERLAP=1): each
        # transformer layer's tensors are one contiguous block of the flat gradient, and its all-reduce can start on the
        # process group's stream as soon as the layer's last gradient kernel is enqueued (reverse layer order).  MEASURED AND
        # LEFT OFF: at 8 GPUs 8.18 ms/step with it, 8.14 without (2 GPUs: 8.05 / 8.04) -- the NCCL kernels take SMs from the
        # persistent one-CTA-per-SM GEMM / attention kernels they run beside, which costs what the overlap saves; the step's
        # 0.3 ms over the single-GPU time is the max over ranks of a synchronised step, not the 10 MB exchange itself.
        pending, done_ranges = [], []
        overlap = (self.world_size > 1 and train_ranges is None and os.environ.get("EAVIT_GRAD_OVERLAP", "0") == "1"
                   and hasattr(rt.encoder, "layer_param_names"))

        def on_layer_done(li):
            for lo, hi in self._layer_ranges(rt, li):
                pending.append(dist.allreduce_sum_async(st.grad[lo:hi]))
                done_ranges.append((lo, hi))
        rt.ac_backward(w["dpol"], w["dv"], backbone=not (train_ranges is not None and self._backbone_frozen),
                       on_layer_done=on_layer_done if overlap else None)
        if side is not None:
            cur.wait_stream(side)
        if train_ranges is not None:                                     # frozen tensors: no gradient (torch leaves .grad = None)
            for lo, hi in self._frozen_ranges:
                call("eavit_zero", st.grad[lo:hi], (hi - lo) * 4)
        call("eavit_add_f32", w["stats"], w["rnd_stats"], w["stats"], 16)
        if self.world_size > 1:                                          # NCCL all-reduce (sum); the mean is applied inside Adam
            if early is not None:
                done_ranges.append(tuple(early))
            for lo, hi in self._complement(done_ranges, st.numel):       # whatever has not been exchanged under the backward
                dist.allreduce_sum_(st.grad[lo:hi])
            for h in pending:
                h.wait()
        if default_config.getboolean("UseGradClipping", fallback=False):
            nrm = torch.zeros(1, dtype=torch.float32, device=rt.device)
            call("eavit_sumsq_f32", st.grad, st.numel, nrm)
            call("eavit_clip_by_norm", st.grad, st.numel, nrm, float(self.max_grad_norm) * self.world_size)
        g = self.optimizer.param_groups[0]
        if apply:
            st.adam_step(g["lr"], 1.0 / self.world_size, g["betas"][0], g["betas"][1], g["eps"], ranges=train_ranges)
            rt.refresh_after_step()
        if stats_out is not None:
            stats_out.copy_(w["stats"])

    def train_model(self, states, target_ext, target_int, y, adv, normalized_extracted_feature_embeddings, old_policy,
                    global_update):
        """agents.py:263-535.  numpy arguments with the reference's dtypes / layouts (flat sample index e*T+t):
        states f32 [N,C,H,W] (/255) or uint8, target_* / adv f64 [N], y int64 [N], next-obs f64 [N,1,H,W]
        (normalised), old_policy f32 [T,E,A].  Mutates parameters and optimiser state; returns None."""
        rt = self.runtime()
        rt.sync()
        R = self.upload_rollout(states, target_ext, target_int, y, adv, normalized_extracted_feature_embeddings, old_policy)
        N = R["states"].shape[0]
        B = self.batch_size
        n_mb = int(N / B)
        steps = self.epoch * n_mb
        # agents.py:336: one torch.rand(B) per minibatch on the CPU generator; drawn up front in the same order
        masks = torch.stack([(torch.rand(B) < self.update_proportion).float() for _ in range(steps)]) if steps else torch.zeros(0, B)
        masks = masks.to(rt.device)
        stats = torch.zeros(max(steps, 1), 16, dtype=torch.float32, device=rt.device)
        sample_range = np.arange(N)                                       # agents.py:270
        k = 0
        for _ in range(self.epoch):
            np.random.shuffle(sample_range)                               # agents.py:276 (MT19937, host)
            perm = torch.from_numpy(sample_range.copy()).to(rt.device)
            for j in range(n_mb):
                self.train_step(R, perm[B * j: B * (j + 1)], masks[k], stats[k])
                k += 1
        self.last_stats = stats[:k]                                       # device tensor; read lazily (no sync here)
        return None

    def stats_summary(self):
        """Mean loss terms of the last update (one D2H read)."""
        if self.last_stats is None or self.last_stats.numel() == 0:
            return {}
        s = self.last_stats.mean(0).cpu().numpy()
        out = dict(actor=s[1], critic_ext=s[2], critic_int=s[3], entropy=s[4], rnd=s[5], approx_kl=s[6], max_kl=s[7], clipfrac=s[8])
        out["loss"] = out["actor"] + 0.5 * (out["critic_ext"] + out["critic_int"]) - self.ent_coef * out["entropy"] + out["rnd"]
        return {k: float(v) for k, v in out.items()}
