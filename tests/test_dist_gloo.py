"""world_size-2 gloo tests (CPU) of the data-parallel host logic: env sharding, gradient mean, moment merges."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as td
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    td.init_process_group("gloo", rank=rank, world_size=world)
    import eavit_b200  # noqa: F401
    from eavit_b200 import dist
    from oracle import oracle as O
    out = {}
    # (a) env sharding: contiguous, disjoint, covering
    out["shard"] = dist.env_shard(8)
    # (b) gradient exchange: all-reduce SUM of the flat buffer, scaled by 1/world inside Adam == mean over shards
    g = torch.full((10,), float(rank + 1))
    dist.allreduce_sum_(g)
    out["grad_mean"] = (g * dist.grad_scale()).tolist()
    # (c) parameter broadcast from rank 0
    p = torch.full((4,), float(rank))
    dist.broadcast_(p, 0)
    out["bcast"] = p.tolist()
    # (d) obs_rms: per-shard partial moments about the shared mean, summed, Chan-merged == oracle over the full batch
    rng = np.random.default_rng(0)
    x = rng.integers(0, 256, (64, 50)).astype(np.float64)          # full batch, identical on both ranks
    lo, hi = (0, 32) if rank == 0 else (32, 64)
    mean0 = torch.zeros(50, dtype=torch.float64)
    var0 = torch.ones(50, dtype=torch.float64)
    xs = torch.from_numpy(x[lo:hi])
    s, qq = (xs - mean0).sum(0), ((xs - mean0) ** 2).sum(0)
    n = torch.tensor([float(hi - lo)], dtype=torch.float64)
    dist.allreduce_sum_(s); dist.allreduce_sum_(qq); dist.allreduce_sum_(n)
    m, v, c = dist.merge_moments(s, qq, float(n.item()), mean0, var0, 1e-4)
    ref = O.RunningMeanStd(shape=(50,))
    ref.update(x)
    out["rms_ok"] = bool(np.allclose(m.numpy(), ref.mean, rtol=1e-12) and np.allclose(v.numpy(), ref.var, rtol=1e-10)
                         and abs(c - ref.count) < 1e-9)
    q.put((rank, out))
    td.destroy_process_group()


def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0]["shard"] == (0, 4) and res[1]["shard"] == (4, 8)
    assert res[0]["grad_mean"] == res[1]["grad_mean"] == [1.5] * 10
    assert res[0]["bcast"] == res[1]["bcast"] == [0.0] * 4
    assert res[0]["rms_ok"] and res[1]["rms_ok"]


def test_multi_shard_oracle_is_mean_of_shard_gradients():
    """the multi-GPU oracle = mean over env shards of the single-process reference gradient.  With equal
    shards the PPO terms are means, so for one shard pair the averaged gradient equals the gradient of the averaged
    loss; checked on a tiny model so it stays fast on CPU."""
    from oracle import oracle as O
    cfg = O.OracleConfig(dim=32, depth=1, heads=2, dim_head=16, mlp_dim=64, patch=12, epoch=1, mini_batch=2, lr=1e-3)
    E, T = 4, 4
    roll = O.synth_rollout(E=E, T=T, seed=5)
    args = O.prepare_update(cfg, T, E, roll, O.RunningMeanStd(shape=(1, 1, 84, 84)), O.RunningMeanStd(),
                            O.RewardForwardFilter(cfg.int_gamma))
    P1, P2 = O.init_params(cfg, seed=3), O.init_params(cfg, seed=3)
    np.random.seed(1); torch.manual_seed(1)
    O.train_model(P1, cfg, *args, n_shards=2, max_steps=1)
    # manual: same permutation / mask, two shard losses averaged, one Adam step
    np.random.seed(1); torch.manual_seed(1)
    N = E * T
    shard, lb = N // 2, (N // cfg.mini_batch) // 2
    perm = np.arange(shard)
    np.random.shuffle(perm)
    mask = (torch.rand(lb) < cfg.update_proportion).float()
    names = O.trainable_names(P2)
    for k in names:
        P2[k].requires_grad_(True)
    opt = torch.optim.Adam([P2[k] for k in names], lr=cfg.lr)
    states, te, ti, y, adv, obs, old = args
    old_flat = torch.tensor(old).permute(1, 0, 2).contiguous().view(-1, cfg.n_actions)
    tot = 0
    for r in range(2):
        idx = torch.from_numpy(perm[:lb] + r * shard)
        loss, _, _ = O.ppo_rnd_loss(P2, cfg, torch.FloatTensor(states)[idx], torch.FloatTensor(te)[idx], torch.FloatTensor(ti)[idx],
                                    torch.LongTensor(y)[idx], torch.FloatTensor(adv)[idx], torch.FloatTensor(obs)[idx], old_flat[idx], mask)
        tot = tot + loss / 2
    tot.backward()
    opt.step()
    for k in names:
        assert torch.allclose(P1[k], P2[k].detach(), rtol=1e-5, atol=1e-7), k
