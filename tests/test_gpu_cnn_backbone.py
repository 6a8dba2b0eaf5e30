"""BASELINE configs[1]: PPO + RND update with the original RND CNN backbone (model.py:110-178, commented out upstream;
restated per SURVEY 8c) -- conv8s4-ReLU-conv4s2-ReLU-conv3s1-ReLU-Flatten-Linear(3136,256)-ReLU-Linear(256,448)-ReLU,
actor 448-448-A, extra_layer 448-448, critics 448-1; policy = actor(x), value = critic(extra_layer(x) + x).
GPU path: eavit_im2col_nchw + tcgen05 GEMM (bf16x3 forward operands) + col2im, the same kernels as the RND towers."""
import numpy as np
import pytest
import torch

from oracle import oracle as O
from test_gpu_model import make_agent, rel, _batch

pytestmark = pytest.mark.gpu
TOL = 1e-2
CFG = O.OracleConfig(impl="cnn", dim=448, lr=1e-3, epoch=2, mini_batch=4)


def test_cnn_forward_u8_and_f32_vs_oracle():
    agent, P = make_agent(CFG, 2, 16)
    rng = np.random.default_rng(11)
    state_u8 = rng.integers(0, 256, (16, 4, 84, 84), dtype=np.uint8)
    state = np.float32(state_u8) / 255.0
    with torch.no_grad():
        pol_o, ve_o, vi_o = O.actor_critic_forward(P, torch.tensor(state), CFG)
        pol, ve, vi = agent.model(torch.tensor(state).cuda())
        pol8, ve8, vi8 = agent.model(torch.tensor(state_u8).cuda())
    assert pol.shape == (16, 18) and ve.shape == (16, 1) and vi.shape == (16, 1)
    assert rel(pol.cpu().numpy(), pol_o.numpy()) < TOL
    assert rel(ve.cpu().numpy(), ve_o.numpy()) < TOL and rel(vi.cpu().numpy(), vi_o.numpy()) < TOL
    # raw frames / 255 in-kernel == np.float32(x) / 255. bit for bit in the patch matrix; the outputs agree to summation-order
    # noise only, because the fully-connected layers accumulate their K splits with red.add
    assert torch.allclose(pol8, pol, rtol=1e-5, atol=1e-7) and torch.allclose(ve8, ve, rtol=1e-5, atol=1e-7)
    np.random.seed(5)
    a, v1, v2, lg = agent.get_action(state)
    np.random.seed(5)
    a_o, v1_o, v2_o, lg_o = O.get_action(P, state, CFG)
    assert rel(lg, lg_o) < TOL and rel(v1, v1_o) < TOL and rel(v2, v2_o) < TOL
    assert (a == a_o).mean() >= 0.9                               # same uniform draw; near-uniform policy: ties within 1e-2 may flip


@pytest.mark.parametrize("B", [16, 256])
def test_cnn_loss_and_gradients_vs_oracle(B):
    """One minibatch of the PPO + RND update (B = 256 is the cfg2 minibatch: 64 envs x 128 steps / 32)."""
    E, T = (2, 16) if B == 16 else (8, 32)
    agent, P = make_agent(CFG, E, T)
    args = _batch(CFG, E, T)
    states, te, ti, y, adv, obs, old = args
    idx = np.random.default_rng(1).permutation(E * T)[:B]
    mask = (np.random.default_rng(2).random(B) < 0.5).astype(np.float32)
    for k in O.trainable_names(P):
        P[k].requires_grad_(True)
    old_flat = torch.tensor(old).permute(1, 0, 2).contiguous().view(-1, CFG.n_actions)
    ti_ = torch.from_numpy(idx)
    loss, terms, _ = O.ppo_rnd_loss(P, CFG, torch.FloatTensor(states)[ti_], torch.FloatTensor(te)[ti_], torch.FloatTensor(ti)[ti_],
                                    torch.LongTensor(y)[ti_], torch.FloatTensor(adv)[ti_], torch.FloatTensor(obs)[ti_], old_flat[ti_],
                                    torch.tensor(mask))
    loss.backward()
    R = agent.upload_rollout(*args)
    stats = torch.zeros(16, device="cuda")
    agent.train_step(R, torch.from_numpy(idx).cuda(), torch.tensor(mask).cuda(), stats, apply=False)
    s = stats.cpu().numpy()
    for k, v in dict(actor=s[1], critic_ext=s[2], critic_int=s[3], entropy=s[4], rnd=s[5]).items():
        assert abs(v - terms[k]) <= TOL * max(abs(terms[k]), 1e-3), (k, v, terms[k])
    st = agent.runtime().store
    parts = {"backbone": ([], []), "heads": ([], []), "rnd": ([], [])}
    for k in O.trainable_names(P):
        assert P[k].grad is not None, k
        part = "rnd" if k.startswith("rnd.") else ("backbone" if k.startswith("model.feature.") else "heads")
        parts[part][0].append(st.g(k).cpu().reshape(-1).numpy()); parts[part][1].append(P[k].grad.reshape(-1).numpy())
    for part, (a, b) in parts.items():
        e = rel(np.concatenate(a), np.concatenate(b))
        assert e < TOL, (part, e)


def test_cnn_train_model_trajectory_and_weights_follow():
    """Whole update through RNDAgent.train_model: same permutation / masks as the oracle, loss terms per step, and the
    bf16x3 operand copies of the backbone follow the optimiser (a stale copy would freeze the forward)."""
    E, T = 2, 16
    agent, P = make_agent(CFG, E, T)
    args = _batch(CFG, E, T)
    np.random.seed(123); torch.manual_seed(123)
    log = O.train_model(P, CFG, *args)
    np.random.seed(123); torch.manual_seed(123)
    agent.train_model(*args, 1)
    stats = agent.last_stats.cpu().numpy()
    assert len(stats) == len(log) == CFG.epoch * CFG.mini_batch
    for i, t in enumerate(log):
        for j, k in ((1, "actor"), (2, "critic_ext"), (3, "critic_int"), (4, "entropy"), (5, "rnd")):
            assert abs(stats[i, j] - t[k]) <= 2e-2 * max(abs(t[k]), 1e-2), (i, k, stats[i, j], t[k])
    # after the update the model must evaluate with the NEW weights (forward vs the oracle at the oracle's new weights)
    sd = agent.state_dict()
    Pn = {k: sd[k].detach().cpu().clone() for k in P}
    x = torch.tensor(args[0][:8])
    with torch.no_grad():
        pol_o, _, _ = O.actor_critic_forward(Pn, x, CFG)
        pol, _, _ = agent.model(x.cuda())
        pol_old, _, _ = O.actor_critic_forward(O.init_params(CFG, seed=7), x, CFG)
    assert rel(pol.cpu().numpy(), pol_o.numpy()) < TOL
    assert rel(pol_old.numpy(), pol_o.numpy()) > 5 * TOL          # the update moved the policy far more than the tolerance
