"""Rollout-side path: RNDAgent.get_action / compute_intrinsic_reward (agents.py:187-218) replayed from
captured CUDA graphs must return exactly what the eager launches return, follow weight updates, and -- with dropout on,
as in the reference's train-mode rollout -- draw fresh masks on every replay."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle as O
from test_gpu_model import CFGS, make_agent, rel

pytestmark = pytest.mark.gpu


def _inputs(E, seed):
    rng = np.random.default_rng(seed)
    states = (rng.integers(0, 256, (E, 4, 84, 84), dtype=np.uint8) / np.float32(255.0)).astype(np.float32)
    obs = rng.normal(0, 1, (E, 1, 84, 84)).clip(-5, 5)
    return states, obs


def _eager(fn, *a):
    os.environ["EAVIT_ROLLOUT_GRAPH"] = "0"
    try:
        return fn(*a)
    finally:
        os.environ["EAVIT_ROLLOUT_GRAPH"] = "1"


@pytest.mark.parametrize("which", ["lucid", "hg"])
def test_graph_replay_equals_eager_and_oracle(which):
    cfg = CFGS[which]
    E = 6
    agent, P = make_agent(cfg, E, 8)
    for seed in (1, 2, 3):                                      # replay with new inputs: the static buffer is refreshed
        states, obs = _inputs(E, seed)
        np.random.seed(5)
        a_g, ve_g, vi_g, pol_g = agent.get_action(states)
        np.random.seed(5)
        a_e, ve_e, vi_e, pol_e = _eager(agent.get_action, states)
        assert np.array_equal(pol_g, pol_e) and np.array_equal(ve_g, ve_e) and np.array_equal(vi_g, vi_e)
        assert np.array_equal(a_g, a_e)                         # same logits + same numpy RNG state -> same actions
        ri_g = agent.compute_intrinsic_reward(obs)
        ri_e = _eager(agent.compute_intrinsic_reward, obs)
        # the towers' fully-connected layers accumulate split-K partial sums atomically: equal up to summation order
        np.testing.assert_allclose(ri_g, ri_e, rtol=2e-5)
        with torch.no_grad():
            pol_o, ve_o, vi_o = O.actor_critic_forward(P, torch.tensor(states), cfg)
            ri_o = O.intrinsic_reward(P, obs)
        assert rel(pol_g, pol_o.numpy()) < 1e-2
        assert rel(ri_g, ri_o) < 1e-2
    assert len(agent._graphs) == 2                              # one graph per call kind, captured once


def test_graph_follows_weight_updates():
    cfg = CFGS["lucid"]
    E, T = 4, 8
    agent, P = make_agent(cfg, E, T)
    states, obs = _inputs(E, 11)
    pol0 = agent.get_action(states)[3]
    ri0 = agent.compute_intrinsic_reward(obs)
    # one fused optimiser step (writes masters + bf16 shadows in place), then load_state_dict (bumps version counters)
    roll = O.synth_rollout(E=E, T=T, seed=21)
    args = O.prepare_update(cfg, T, E, roll, O.RunningMeanStd(shape=(1, 1, 84, 84)), O.RunningMeanStd(),
                            O.RewardForwardFilter(cfg.int_gamma))
    R = agent.upload_rollout(*args)
    B = 8
    agent.train_step(R, torch.arange(B, device="cuda"), torch.ones(B, device="cuda"))
    pol1 = agent.get_action(states)[3]
    ri1 = agent.compute_intrinsic_reward(obs)
    assert not np.array_equal(pol0, pol1) and not np.allclose(ri0, ri1, rtol=1e-4)
    assert np.array_equal(pol1, _eager(agent.get_action, states)[3])
    np.testing.assert_allclose(ri1, _eager(agent.compute_intrinsic_reward, obs), rtol=2e-5)
    agent.load_state_dict({k: v.clone() for k, v in P.items()}, strict=True)
    assert np.array_equal(agent.get_action(states)[3], pol0)     # sync_if_changed saw the load
    np.testing.assert_allclose(agent.compute_intrinsic_reward(obs), ri0, rtol=2e-5)


def test_graph_replays_draw_fresh_dropout_masks():
    cfg = CFGS["lucid"]
    E = 8
    agent, P = make_agent(cfg, E, 8, ViTlucidrains_dropout=0.1, ViTlucidrains_emb_dropout=0.1)
    agent.set_mode("train")
    states, _ = _inputs(E, 3)
    pols = np.stack([agent.get_action(states)[3] for _ in range(24)])
    assert len(agent._graphs) == 1
    # every replay is a different mask draw ...
    assert all(not np.array_equal(pols[i], pols[j]) for i in range(6) for j in range(i))
    # ... around the dropout-free forward: unbiased masks, so the mean over draws approaches it
    agent.set_mode("eval")
    clean = agent.get_action(states)[3]
    assert rel(pols.mean(0), clean) < 0.5 * rel(pols[0], clean) + 1e-3
    eager = np.stack([_eager(agent.get_action, states)[3] for _ in range(2)])
    assert np.array_equal(eager[0], eager[1]) and np.array_equal(eager[0], clean)     # eval mode: no dropout, no epoch effect


def test_device_rollout_buffer_matches_reference_glue():
    """DeviceRollout (per-step env-major writes + finish()) == the oracle's restatement of train.py:707-779 / :855 fed with
    the same step-major buffers; then the tuple drives train_model directly (CUDA tensors, uint8 frames)."""
    import eavit_b200  # noqa
    from eavit_b200 import config, rollout, utils
    cfg = CFGS["lucid"]
    E, T, A = 4, 8, cfg.n_actions
    roll = O.synth_rollout(E=E, T=T, seed=33)
    rng = np.random.default_rng(5)
    warm = rng.integers(0, 256, (40, 1, 84, 84)).astype(np.float64)      # initial observation statistics (train.py:127-180)
    prev_r = rng.random((E, T)).astype(np.float32)                        # an earlier update: the filter state persists
    # ---- oracle
    o_obs, o_rrm, o_f = O.RunningMeanStd(shape=(1, 1, 84, 84)), O.RunningMeanStd(), O.RewardForwardFilter(cfg.int_gamma)
    o_obs.update(warm)
    O.normalize_int_reward(prev_r, o_f, o_rrm)
    ref = O.prepare_update(cfg, T, E, roll, o_obs, o_rrm, o_f)
    # ---- device
    config.load_config(None, TrainMethod="original_RND")
    d_obs, d_rrm = utils.RunningMeanStd(shape=(1, 1, 84, 84), usage="obs_rms"), utils.RunningMeanStd(usage="reward_rms")
    d_f = utils.RewardForwardFilter(cfg.int_gamma)
    d_obs.update(warm)
    mom = d_f.filter_rollout(torch.tensor(prev_r, device="cuda")).cpu().numpy()
    d_rrm.update_from_moments(float(mom[0]), float(mom[1]), int(mom[2]))
    buf = rollout.DeviceRollout(E, T, A)
    ve, vi = roll["total_ext_values"].reshape(T + 1, E), roll["total_int_values"].reshape(T + 1, E)
    for t in range(T):
        sl = slice(t * E, (t + 1) * E)
        buf.add(t, roll["total_state"][sl], roll["total_next_obs"][sl], roll["total_reward"][sl], roll["total_done"][sl],
                roll["total_action"][sl], ve[t], vi[t], roll["total_policy"][sl], roll["total_int_reward"][sl])
    buf.add_last_values(ve[T], vi[T])
    got = buf.finish(d_obs, d_rrm, d_f, cfg.gamma, cfg.int_gamma, cfg.lam, cfg.ext_coef, cfg.int_coef)
    states, te, ti, y, adv, obs, old = [g.cpu().numpy() for g in got]
    r_states, r_te, r_ti, r_y, r_adv, r_obs, r_old = ref
    assert states.dtype == np.uint8 and np.array_equal(np.float32(states) / np.float32(255.0), r_states)     # layout e*T+t, exact
    assert np.array_equal(y, r_y)
    assert np.array_equal(old, r_old.transpose(1, 0, 2).reshape(E * T, A))                                    # agents.py:301
    assert np.array_equal(te, r_te)                                       # extrinsic GAE: float64, bit-exact
    np.testing.assert_allclose(ti, r_ti, rtol=2e-5)                       # intrinsic: through the float32 reward moments
    np.testing.assert_allclose(adv, r_adv, rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(obs, np.float32(r_obs), rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(d_rrm.var, o_rrm.var, rtol=1e-5)
    assert d_rrm.count == o_rrm.count and np.allclose(d_f.rewems, o_f.rewems, rtol=1e-6)
    np.testing.assert_allclose(d_obs.mean, o_obs.mean, rtol=1e-12, atol=1e-12)
    # ---- and the tuple is a valid train_model argument list (device tensors, no host copies)
    agent, P = make_agent(cfg, E, T)
    before = agent.state_dict()["model.actor.2.weight"].clone()
    np.random.seed(0); torch.manual_seed(0)
    agent.train_model(*got, 0)
    torch.cuda.synchronize()
    assert not torch.equal(before, agent.state_dict()["model.actor.2.weight"])
    s = agent.stats_summary()
    assert np.isfinite(s["loss"])
