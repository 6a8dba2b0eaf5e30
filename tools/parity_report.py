"""Measured values behind every assert of tests/test_gpu_model.py that is looser than the north-star 1e-2:
value heads vs their bf16 floor, per-tensor gradient errors, trajectory loss terms, update direction.  Run on a B200:

    python tools/parity_report.py > profiles/r2_parity_measured.txt
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from oracle import oracle as O
from test_gpu_model import CFGS, make_agent, rel, bf16_floor, _batch

G = {w: np.load(os.path.join(ROOT, "tests", "golden", f"golden_{w}.npz")) for w in ("lucid", "cls", "hg")}
print("== forward vs reference golden (norm-wise relative error) and the ideal-bf16 floor of the same weights")
for which in ("lucid", "cls", "hg"):
    agent, P = make_agent(CFGS[which], 2, 16)
    rng = np.random.default_rng(11)
    state = np.float32(rng.integers(0, 256, (16, 4, 84, 84), dtype=np.uint8)) / 255.0
    with torch.no_grad():
        pol, ve, vi = agent.model(torch.tensor(state).cuda())
    fl = bf16_floor(CFGS[which], P, torch.tensor(state))
    print(f"{which:6s} policy {rel(pol.cpu().numpy(), G[which]['fwd_policy']):.2e} (floor {fl[0]:.2e})   value_ext {rel(ve.cpu().numpy(), G[which]['fwd_value_ext']):.2e} "
          f"(floor {fl[1]:.2e})   value_int {rel(vi.cpu().numpy(), G[which]['fwd_value_int']):.2e} (floor {fl[2]:.2e})")
print("== one minibatch: total / worst per-tensor gradient error vs the oracle (tensors carrying > 2 % of the gradient norm)")
for which in ("lucid", "cls", "hg"):
    cfg = CFGS[which]
    E, T, B = 2, 16, 16
    agent, P = make_agent(cfg, E, T)
    args = _batch(cfg, E, T)
    states, te, ti, y, adv, obs, old = args
    idx = np.random.default_rng(1).permutation(E * T)[:B]
    mask = (np.random.default_rng(2).random(B) < 0.5).astype(np.float32)
    for k in O.trainable_names(P):
        P[k].requires_grad_(True)
    old_flat = torch.tensor(old).permute(1, 0, 2).contiguous().view(-1, cfg.n_actions)
    ti_ = torch.from_numpy(idx)
    loss, terms, _ = O.ppo_rnd_loss(P, cfg, torch.FloatTensor(states)[ti_], torch.FloatTensor(te)[ti_], torch.FloatTensor(ti)[ti_],
                                    torch.LongTensor(y)[ti_], torch.FloatTensor(adv)[ti_], torch.FloatTensor(obs)[ti_], old_flat[ti_],
                                    torch.tensor(mask))
    loss.backward()
    R = agent.upload_rollout(*args)
    agent.train_step(R, torch.from_numpy(idx).cuda(), torch.tensor(mask).cuda(), None, apply=False)
    st = agent.runtime().store
    worst, ref, got = {}, [], []
    for k in O.trainable_names(P):
        if P[k].grad is None or k.endswith("attention.key.bias"):
            continue
        a, b = st.g(k).cpu().reshape(-1).numpy(), P[k].grad.reshape(-1).numpy()
        ref.append(b); got.append(a); worst[k] = (rel(a, b), float(np.linalg.norm(b)))
    tn = float(np.linalg.norm(np.concatenate(ref)))
    big = sorted(((v[0], k) for k, v in worst.items() if v[1] > 0.02 * tn), reverse=True)[:4]
    small = sorted(((v[0], v[1] / tn, k) for k, v in worst.items() if v[1] <= 0.02 * tn), reverse=True)[:3]
    print(f"{which:6s} total {rel(np.concatenate(got), np.concatenate(ref)):.2e}   worst of the tensors > 2 % of the norm: " +
          ", ".join(f"{k.split('feature.')[-1]} {e:.2e}" for e, k in big))
    print("         worst of the small tensors (error, share of the gradient norm): " + ", ".join(f"{k.split('feature.')[-1]} {e:.2e} ({s:.1e})" for e, s, k in small))
print("== whole update through train_model (lucid, 8 optimiser steps at lr 1e-3): per-step loss terms and the parameter update")
cfg = CFGS["lucid"]
agent, P = make_agent(cfg, 2, 16)
args = _batch(cfg, 2, 16)
P0 = {k: v.clone() for k, v in P.items()}
np.random.seed(123); torch.manual_seed(123)
log = O.train_model(P, cfg, *args)
np.random.seed(123); torch.manual_seed(123)
agent.train_model(*args, 1)
stats = agent.last_stats.cpu().numpy()
dev = {k: 0.0 for k in ("actor", "critic_ext", "critic_int", "entropy", "rnd")}
for i, t in enumerate(log):
    for j, k in ((1, "actor"), (2, "critic_ext"), (3, "critic_int"), (4, "entropy"), (5, "rnd")):
        dev[k] = max(dev[k], abs(stats[i, j] - t[k]) / max(abs(t[k]), 1e-2))
print("max relative deviation of a loss term over the 8 steps:", {k: f"{v:.2e}" for k, v in dev.items()})
sd = agent.state_dict()
num = den1 = den2 = 0.0
for k in O.trainable_names(P):
    d_ref = (P[k].detach() - P0[k]).reshape(-1).double().numpy()
    d_got = (sd[k].cpu() - P0[k]).reshape(-1).double().numpy()
    num += float(d_ref @ d_got); den1 += float(d_ref @ d_ref); den2 += float(d_got @ d_got)
print(f"update direction cosine {num / np.sqrt(den1 * den2):.4f}, update norm ratio {np.sqrt(den2 / den1):.4f}  "
      "(Adam's first steps move every weight by ~lr * sign(g): elements whose gradient is rounding noise take a random sign)")
