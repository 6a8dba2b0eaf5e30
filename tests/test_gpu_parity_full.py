"""Parity at the shapes that are actually benched / shipped ("close the parity gaps"):

  * cfg3 at the benched minibatch: 512 samples (T = 201 216 tokens -- the persistent-grid, split-K and packed paths the
    16-sample tests never reach), loss terms + flat gradient vs the CPU oracle;
  * the HF-style ViT at its shipped size (1024 / 12 layers / 16 heads / 3072), forward + every gradient vs the oracle
    (which ``tests/test_oracle_golden.py`` pins to the unmodified reference at the same size and inputs) and vs the
    reference golden itself;
  * a 2-rank NCCL update vs the shard oracle ``O.train_model(n_shards=2)`` (needs 2 GPUs, skipped and said so otherwise).

Tolerance: north_star's 1e-2 norm-wise relative for bf16 outputs / gradients."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import oracle as O
from test_gpu_model import make_agent, rel, bf16_floor

pytestmark = pytest.mark.gpu
TOL = 1e-2
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cfg3_benched_minibatch_512_loss_and_flat_gradient_vs_oracle():
    cfg = O.OracleConfig()                     # the cfg3 model: dim 256, depth 3, 8 x 32 heads, mlp 1024, patch 6
    E, T, B = 8, 64, 512
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    agent, P = make_agent(cfg, E, T)
    roll = O.synth_rollout(E=E, T=T, seed=77)
    args = O.prepare_update(cfg, T, E, roll, O.RunningMeanStd(shape=(1, 1, 84, 84)), O.RunningMeanStd(), O.RewardForwardFilter(cfg.int_gamma))
    states, te, ti, y, adv, obs, old = args
    idx = np.random.default_rng(1).permutation(E * T)[:B]
    mask = (np.random.default_rng(2).random(B) < cfg.update_proportion).astype(np.float32)
    for k in O.trainable_names(P):
        P[k].requires_grad_(True)
    old_flat = torch.tensor(old).permute(1, 0, 2).contiguous().view(-1, cfg.n_actions)
    ti_ = torch.from_numpy(idx)
    terms = O.ppo_rnd_backward_chunked(P, cfg, torch.FloatTensor(states)[ti_], torch.FloatTensor(te)[ti_], torch.FloatTensor(ti)[ti_],
                                       torch.LongTensor(y)[ti_], torch.FloatTensor(adv)[ti_], torch.FloatTensor(obs)[ti_],
                                       old_flat[ti_], torch.tensor(mask), chunk=64)
    R = agent.upload_rollout(*args)
    stats = torch.zeros(16, device="cuda")
    agent.train_step(R, torch.from_numpy(idx).cuda(), torch.tensor(mask).cuda(), stats, apply=False)
    s = stats.cpu().numpy()
    got = dict(actor=s[1], critic_ext=s[2], critic_int=s[3], entropy=s[4], rnd=s[5])
    for k, v in got.items():
        assert abs(v - terms[k]) <= TOL * max(abs(terms[k]), 1e-3), (k, v, terms[k])
    st = agent.runtime().store
    assert st.numel > 4_000_000
    ref, mine, parts = [], [], {"vit": ([], []), "heads": ([], []), "rnd": ([], [])}
    for k in O.trainable_names(P):
        if P[k].grad is None:
            assert float(st.g(k).abs().max()) == 0.0, k
            continue
        a, b = st.g(k).cpu().reshape(-1).numpy(), P[k].grad.reshape(-1).numpy()
        mine.append(a); ref.append(b)
        part = "rnd" if k.startswith("rnd.") else ("vit" if k.startswith("model.feature.") else "heads")
        parts[part][0].append(a); parts[part][1].append(b)
    tot = rel(np.concatenate(mine), np.concatenate(ref))
    assert tot < TOL, tot
    for part, (a, b) in parts.items():         # each sub-network on its own, so a small one cannot hide behind a large one
        e = rel(np.concatenate(a), np.concatenate(b))
        assert e < TOL, (part, e)


def _hg_full():
    from test_oracle_golden import hg_full_cfg, hg_full_inputs
    return hg_full_cfg(), hg_full_inputs()


def test_hg_shipped_size_forward_and_gradients(golden_dir):
    """ViT_ExplorativeAttn at 1024 / 12 L / 16 h / 3072 (vit_hg.py:277-374, model.py:200-220, the shipped ViTHG_* keys)."""
    cfg, (state, w) = _hg_full()
    G = np.load(os.path.join(golden_dir, "golden_hg_full.npz"))
    agent, P = make_agent(cfg, 2, 4)
    x = torch.tensor(state)
    # forward vs the REFERENCE's own outputs
    with torch.no_grad():
        pol, ve, vi = agent.model(x.cuda())
    floor = bf16_floor(cfg, P, x)
    assert rel(pol.cpu().numpy(), G["fwd_policy"]) < TOL
    # values: 1e-2, or the ideal-bf16 floor of this weight set where cancellation in the 1024-term value heads lifts it
    # above 1e-2 (see test_gpu_model.bf16_floor; the measured figures are printed on failure)
    ev, ei = rel(ve.cpu().numpy(), G["fwd_value_ext"]), rel(vi.cpu().numpy(), G["fwd_value_int"])
    assert ev < max(TOL, 1.25 * floor[1]) and ei < max(TOL, 1.25 * floor[2]), (ev, ei, floor)
    # backward: every parameter gradient vs the oracle (full tensors), and the reference's digests as a cross-check
    for k in P:
        if k.startswith("model."):
            P[k].requires_grad_(True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    pol_o, ve_o, vi_o = O.actor_critic_forward(P, x, cfg)
    ((pol_o * torch.tensor(w)).sum() + 3.0 * ve_o.sum() + 2.0 * vi_o.sum()).backward()
    agent.optimizer.zero_grad()
    pol, ve, vi = agent.model(x.cuda())
    ((pol * torch.tensor(w).cuda()).sum() + 3.0 * ve.sum() + 2.0 * vi.sum()).backward()
    ref, got = [], []
    worst = {}
    gsq_ref = gsq_got = 0.0
    for name, p in agent.named_parameters():
        if not name.startswith("model.") or P[name].grad is None or name.endswith("attention.key.bias"):
            continue
        a, b = p.grad.detach().cpu().reshape(-1).numpy(), P[name].grad.reshape(-1).numpy()
        got.append(a); ref.append(b)
        worst[name] = rel(a, b)
        gsq_ref += float(G["grad/" + name][0]) ** 2
        gsq_got += float(np.linalg.norm(a.astype(np.float64))) ** 2
    ref_all = np.concatenate(ref)
    tot = rel(np.concatenate(got), ref_all)
    assert tot < TOL, (tot, sorted(worst.items(), key=lambda kv: -kv[1])[:8])
    # the reference's own gradient (digest norms of `make_golden.py hg_full`): total norm, and every tensor that carries
    # more than 2 % of it
    assert abs(np.sqrt(gsq_got) - np.sqrt(gsq_ref)) < TOL * np.sqrt(gsq_ref), (gsq_got, gsq_ref)
    tn = float(np.linalg.norm(ref_all))
    for name, p in agent.named_parameters():
        if name in worst and float(G["grad/" + name][0]) > 0.02 * tn:
            gn = float(p.grad.detach().double().norm())
            assert abs(gn - float(G["grad/" + name][0])) < 3 * TOL * float(G["grad/" + name][0]), (name, gn, float(G["grad/" + name][0]))
    bad = {k: v for k, v in worst.items() if v > 5 * TOL and float(P[k].grad.norm()) > 0.02 * tn}
    assert not bad, sorted(bad.items(), key=lambda kv: -kv[1])[:8]


_RANK_SCRIPT = r'''
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.environ["EAVIT_ROOT"]); sys.path.insert(0, os.path.join(os.environ["EAVIT_ROOT"], "tests"))
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", rank))
from oracle import oracle as O
from test_gpu_model import make_agent
cfg = O.OracleConfig(lr=1e-3, epoch=1, mini_batch=2)
E, T = 4, 8                                      # global; every rank owns E / world envs (contiguous env shards)
roll = O.synth_rollout(E=E, T=T, seed=21)
args = O.prepare_update(cfg, T, E, roll, O.RunningMeanStd(shape=(1, 1, 84, 84)), O.RunningMeanStd(), O.RewardForwardFilter(cfg.int_gamma))
states, te, ti, y, adv, obs, old = args
Es = E // world
lo, hi = rank * Es * T, (rank + 1) * Es * T      # flat sample index = e * T + t  -> contiguous per env shard
agent, P = make_agent(cfg, Es, T)
# (a) ONE minibatch, gradient only: after the all-reduce every rank holds the SUM of the shard gradients
local = (states[lo:hi], te[lo:hi], ti[lo:hi], y[lo:hi], adv[lo:hi], obs[lo:hi], old[:, rank * Es:(rank + 1) * Es])
R = agent.upload_rollout(*local)
agent.runtime()
Bl = Es * T // cfg.mini_batch
idx = torch.arange(Bl, device="cuda") * 2 % (Es * T)
mask = torch.tensor((np.arange(Bl) % 3 != 0).astype(np.float32)).cuda()
stats = torch.zeros(16, device="cuda")
agent.train_step(R, idx, mask, stats, apply=False)
torch.cuda.synchronize()
st = agent.runtime().store
grad_mean = {k: (st.g(k) / world).cpu() for k in st.shapes}
stats_all = [torch.zeros_like(stats) for _ in range(world)]
torch.distributed.all_gather(stats_all, stats)
# (b) a whole update through the reference-facing call
np.random.seed(123); torch.manual_seed(123)      # every rank holds the same seeds (train.py:52)
agent.train_model(states[lo:hi], te[lo:hi], ti[lo:hi], y[lo:hi], adv[lo:hi], obs[lo:hi], old[:, rank * Es:(rank + 1) * Es], 1)
torch.cuda.synchronize()
sd = {k: v.detach().cpu() for k, v in agent.state_dict().items()}
if rank == 0:
    torch.save({"sd": sd, "stats": agent.last_stats.cpu(), "grad_mean": grad_mean, "step_stats": torch.stack(stats_all).cpu(),
                "idx": idx.cpu(), "mask": mask.cpu()}, os.environ["EAVIT_OUT"])
flat = agent.runtime().store.flat
a, b = flat.clone(), flat.clone()
torch.distributed.all_reduce(a, op=torch.distributed.ReduceOp.MIN); torch.distributed.all_reduce(b, op=torch.distributed.ReduceOp.MAX)
assert torch.equal(a, b), "ranks diverged"
torch.distributed.barrier(); torch.distributed.destroy_process_group()
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs: the N-rank NCCL update vs the shard oracle runs under "
                                                          "`gpurun --gpus 2` (profiles/r2_two_rank_nccl_vs_shard_oracle.txt)")
def test_two_rank_nccl_update_matches_shard_oracle(tmp_path):
    """What replaces train.py:243 / :854 (DDP wrapper + gradient averaging): two NCCL ranks, each training on its env shard
    with an all-reduced gradient, end with the weights of ``O.train_model(n_shards=2)`` (same permutation and masks on
    every rank, shard-mean gradient, ONE Adam step per minibatch)."""
    cfg = O.OracleConfig(lr=1e-3, epoch=1, mini_batch=2)
    E, T = 4, 8
    out = str(tmp_path / "rank0.pt")
    script = tmp_path / "rank.py"
    script.write_text(_RANK_SCRIPT)
    env = dict(os.environ, EAVIT_ROOT=ROOT, EAVIT_OUT=out, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29517", str(script)], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    got = torch.load(out)
    roll = O.synth_rollout(E=E, T=T, seed=21)
    args = O.prepare_update(cfg, T, E, roll, O.RunningMeanStd(shape=(1, 1, 84, 84)), O.RunningMeanStd(), O.RewardForwardFilter(cfg.int_gamma))
    # (a) shard-mean gradient of one minibatch: the data-parallel semantics themselves, at the gradient tolerance
    states, te, ti, y, adv, obs, old = args
    P = O.init_params(cfg, seed=7)
    for k in O.trainable_names(P):
        P[k].requires_grad_(True)
    old_flat = torch.tensor(old).permute(1, 0, 2).contiguous().view(-1, cfg.n_actions)
    shard = E * T // 2
    terms_mean = None
    for r in range(2):
        ii = got["idx"] + r * shard
        loss, terms, _ = O.ppo_rnd_loss(P, cfg, torch.FloatTensor(states)[ii], torch.FloatTensor(te)[ii], torch.FloatTensor(ti)[ii],
                                        torch.LongTensor(y)[ii], torch.FloatTensor(adv)[ii], torch.FloatTensor(obs)[ii], old_flat[ii],
                                        got["mask"])
        (loss / 2).backward()
        for j, k in ((1, "actor"), (2, "critic_ext"), (3, "critic_int"), (4, "entropy"), (5, "rnd")):
            v = float(got["step_stats"][r, j])                       # every rank reports the loss terms of ITS shard
            assert abs(v - terms[k]) <= TOL * max(abs(terms[k]), 1e-3), (r, k, v, terms[k])
    ref = np.concatenate([P[k].grad.reshape(-1).numpy() for k in O.trainable_names(P) if P[k].grad is not None])
    mine = np.concatenate([got["grad_mean"][k].reshape(-1).numpy() for k in O.trainable_names(P) if P[k].grad is not None])
    assert rel(mine, ref) < TOL, rel(mine, ref)
    # (b) the whole update
    P = O.init_params(cfg, seed=7)
    P0 = {k: v.clone() for k, v in P.items()}
    np.random.seed(123); torch.manual_seed(123)
    log = O.train_model(P, cfg, *args, n_shards=2)
    # Adam's first steps move every weight by ~lr * sign(g): compare the UPDATE (direction and size), tensor by tensor
    num = den1 = den2 = 0.0
    for k in O.trainable_names(P):
        d_ref = (P[k].detach() - P0[k]).reshape(-1).double().numpy()
        d_got = (got["sd"][k] - P0[k]).reshape(-1).double().numpy()
        num += float(d_ref @ d_got); den1 += float(d_ref @ d_ref); den2 += float(d_got @ d_got)
    assert den1 > 0
    cos = num / np.sqrt(den1 * den2)
    assert cos > 0.97, cos
    assert abs(np.sqrt(den2 / den1) - 1) < 0.03
    # per-step loss terms: rank 0's own shard vs the oracle's shard-0 ... the oracle logs the shard MEAN; rank 0 logs its
    # shard only, so compare the first step of a 1-minibatch update through the weights instead (above) and check counts
    assert len(log) == got["stats"].shape[0]
