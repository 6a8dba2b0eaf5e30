"""Edge shapes of the drop-in calls: one env, odd batch sizes (rows that are no multiple of any tile), a rollout whose
sample count is not a multiple of the minibatch size (the reference drops the remainder: int(N / batch), agents.py:284),
and one-step rollouts through the numerics kernels."""
import numpy as np
import pytest
import torch

from oracle import oracle as O
from test_gpu_model import CFGS, make_agent, rel

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("E", [1, 3, 7])
def test_rollout_calls_with_odd_env_counts(E):
    cfg = CFGS["lucid"]
    agent, P = make_agent(cfg, E, 8)
    rng = np.random.default_rng(E)
    states = (rng.integers(0, 256, (E, 4, 84, 84), dtype=np.uint8) / np.float32(255.0)).astype(np.float32)
    obs = rng.normal(0, 1, (E, 1, 84, 84)).clip(-5, 5)
    u = rng.random(E)
    np.random.seed(3)
    action, ve, vi, pol = agent.get_action(states)
    assert action.shape == (E,) and action.dtype == np.int64 and pol.shape == (E, cfg.n_actions)
    a_o, ve_o, vi_o, pol_o = O.get_action(P, states, cfg, u=u)
    assert rel(pol, pol_o) < 1e-2
    # values are O(0.05) sums of 256 terms: compare on the scale of the logits' accuracy (see bf16_floor in test_gpu_model)
    assert np.abs(np.atleast_1d(ve) - np.atleast_1d(ve_o)).max() < 5e-3 and np.abs(np.atleast_1d(vi) - np.atleast_1d(vi_o)).max() < 5e-3
    ri = agent.compute_intrinsic_reward(obs)
    assert ri.shape == (E,) and ri.dtype == np.float32
    assert rel(ri, O.intrinsic_reward(P, obs)) < 1e-2


def test_train_model_drops_the_remainder_like_the_reference():
    """N = 15 samples, batch 4 -> 3 minibatches per epoch, 3 samples never used in an epoch (agents.py:284)."""
    cfg = O.OracleConfig(lr=1e-3, epoch=2, mini_batch=3)        # oracle: batch = N // mini_batch = 5 -> use N = 15, batch 5
    E, T = 3, 5
    agent, P = make_agent(cfg, E, T)
    agent.batch_size = 4                                         # not a divisor of N = 15
    roll = O.synth_rollout(E=E, T=T, seed=9)
    args = O.prepare_update(cfg, T, E, roll, O.RunningMeanStd(shape=(1, 1, 84, 84)), O.RunningMeanStd(), O.RewardForwardFilter(cfg.int_gamma))
    np.random.seed(11); torch.manual_seed(11)
    agent.train_model(*args, 0)
    stats = agent.last_stats.cpu().numpy()
    assert len(stats) == cfg.epoch * (15 // 4)
    assert np.isfinite(stats).all()
    # the same permutation / mask stream drives an oracle with the same batch size: loss terms step by step
    perm = np.arange(15)
    np.random.seed(11); torch.manual_seed(11)
    masks = [(torch.rand(4) < 0.25).float() for _ in range(len(stats))]
    states, te, ti, y, adv, obs, old = args
    old_flat = torch.tensor(old).permute(1, 0, 2).contiguous().view(-1, cfg.n_actions)
    for k in O.trainable_names(P):
        P[k].requires_grad_(True)
    opt = torch.optim.Adam([P[k] for k in O.trainable_names(P)], lr=cfg.lr)
    k = 0
    for _ in range(cfg.epoch):
        np.random.shuffle(perm)
        for j in range(15 // 4):
            idx = torch.from_numpy(perm[4 * j: 4 * (j + 1)].copy())
            opt.zero_grad()
            loss, terms, _ = O.ppo_rnd_loss(P, cfg, torch.FloatTensor(states)[idx], torch.FloatTensor(te)[idx], torch.FloatTensor(ti)[idx],
                                            torch.LongTensor(y)[idx], torch.FloatTensor(adv)[idx], torch.FloatTensor(obs)[idx],
                                            old_flat[idx], masks[k])
            loss.backward()
            opt.step()
            for col, name in ((1, "actor"), (2, "critic_ext"), (3, "critic_int"), (4, "entropy"), (5, "rnd")):
                assert abs(stats[k, col] - float(terms[name])) <= 3e-2 * max(abs(float(terms[name])), 1e-2), (k, name, stats[k, col], float(terms[name]))
            k += 1


def test_one_step_rollout_numerics():
    """T = 1 and E = 1 through GAE, the reward filter and the device rollout buffer."""
    import eavit_b200  # noqa
    from eavit_b200 import config, rollout, utils
    config.load_config(None, TrainMethod="original_RND")
    cfg = O.OracleConfig()
    for E, T in ((1, 1), (5, 1), (1, 6)):
        roll = O.synth_rollout(E=E, T=T, seed=E * 10 + T)
        ref = O.prepare_update(cfg, T, E, roll, O.RunningMeanStd(shape=(1, 1, 84, 84)), O.RunningMeanStd(), O.RewardForwardFilter(cfg.int_gamma))
        buf = rollout.DeviceRollout(E, T, cfg.n_actions)
        ve, vi = roll["total_ext_values"].reshape(T + 1, E), roll["total_int_values"].reshape(T + 1, E)
        for t in range(T):
            sl = slice(t * E, (t + 1) * E)
            buf.add(t, roll["total_state"][sl], roll["total_next_obs"][sl], roll["total_reward"][sl], roll["total_done"][sl],
                    roll["total_action"][sl], ve[t], vi[t], roll["total_policy"][sl], roll["total_int_reward"][sl])
        buf.add_last_values(ve[T], vi[T])
        got = buf.finish(utils.RunningMeanStd(shape=(1, 1, 84, 84), usage="obs_rms"), utils.RunningMeanStd(usage="reward_rms"),
                         utils.RewardForwardFilter(cfg.int_gamma), cfg.gamma, cfg.int_gamma, cfg.lam, cfg.ext_coef, cfg.int_coef)
        states, te, ti, y, adv, obs, old = [g.cpu().numpy() for g in got]
        assert np.array_equal(te, ref[1]) and np.array_equal(y, ref[3])
        np.testing.assert_allclose(ti, ref[2], rtol=2e-5, atol=1e-6)
        np.testing.assert_allclose(adv, ref[4], rtol=2e-5, atol=1e-6)
        if E * T > 1:                                            # a single frame has zero variance: the reference divides by zero too
            np.testing.assert_allclose(obs, np.float32(ref[5]), rtol=1e-5, atol=1e-5)
        # numpy-facing GAE entry point on the same shapes
        r = roll["total_reward"].reshape(T, E).transpose().clip(-1, 1)
        d = roll["total_done"].reshape(T, E).transpose()
        tgt, a2 = utils.make_train_data(r, d, ve.transpose().copy(), cfg.gamma, T, E)
        o_t, o_a = O.make_train_data(r, d, ve.transpose().copy(), cfg.gamma, T, E, cfg.lam)
        assert np.array_equal(tgt, o_t) and np.array_equal(a2, o_a)
