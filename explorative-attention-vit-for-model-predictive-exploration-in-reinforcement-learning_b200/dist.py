"""Data-parallel plumbing: one process per GPU, torch.distributed (NCCL over NVLink on the box, gloo in CPU tests).

The hot path shards by environment: every per-sample / per-env computation is independent, so there
is no data-path collective.  The only exchange steps are
  (1) the gradient all-reduce (mean) once per optimiser step over the agent's ONE flat gradient buffer,
  (2) the obs_rms moments (sum, sum of squares about the current mean, count) once per update,
  (3) the reward-filter moments (sum, sumsq, n) once per update.
The reference wraps the agent in DDP but never arms the reducer; this implements the intended
semantics.  All helpers are no-ops when torch.distributed is not initialised.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as td


def is_dist() -> bool:
    return td.is_available() and td.is_initialized() and td.get_world_size() > 1


def world() -> Tuple[int, int]:
    return (td.get_world_size(), td.get_rank()) if is_dist() else (1, 0)


def env_shard(num_env_total: int) -> Tuple[int, int]:
    """Contiguous env range [lo, hi) owned by this rank (dist_utils.py:104-109 partitions envs by rank)."""
    w, r = world()
    assert num_env_total % w == 0, "envs must divide evenly over ranks"
    per = num_env_total // w
    return r * per, (r + 1) * per


def broadcast_(t: torch.Tensor, src: int = 0):
    if is_dist():
        td.broadcast(t, src)
    return t


def allreduce_sum_(t: torch.Tensor):
    if is_dist():
        td.all_reduce(t, op=td.ReduceOp.SUM)
    return t


def allreduce_sum_async(t: torch.Tensor):
    """Start the sum all-reduce of ``t`` on the process group's own stream, ordered after everything enqueued so far on the
    current stream; returns the Work handle (``.wait()`` makes the current stream wait for the result), None when not
    distributed."""
    if is_dist():
        return td.all_reduce(t, op=td.ReduceOp.SUM, async_op=True)
    return None


def grad_scale() -> float:
    """Multiplier that turns the all-reduced gradient SUM into the mean over ranks (applied inside Adam)."""
    return 1.0 / world()[0]


def merge_moments(total_sum: torch.Tensor, total_sumsq: torch.Tensor, total_count: float, mean: torch.Tensor,
                  var: torch.Tensor, count: float):
    """Chan merge (utils.py:101-115) of batch moments given as sums about the CURRENT mean `mean`.
    Pure tensor arithmetic (float64) -- the device path uses eavit_rms_merge; this is its host mirror for tests."""
    ds = total_sum / total_count
    b_var = torch.clamp(total_sumsq / total_count - ds * ds, min=0.0)
    tot = count + total_count
    new_mean = mean + ds * total_count / tot
    m2 = var * count + b_var * total_count + ds * ds * count * total_count / tot
    return new_mean, m2 / tot, tot
