"""Pins the CPU oracle (oracle/oracle.py) against fixtures produced by the unmodified reference
(tests/golden/make_golden.py, run in the build container).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle as O


def digest(a):
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    idx = np.linspace(0, a.size - 1, 64).astype(np.int64)
    return np.concatenate([[np.sqrt((a * a).sum()), a.sum()], a[idx]])


CFGS = {
    "lucid": O.OracleConfig(lr=1e-3, epoch=2, mini_batch=4),
    "cls": O.OracleConfig(use_explorative=False),
    "hg": O.OracleConfig(impl="hg", patch=12, dim=128, depth=2, heads=2, dim_head=64, mlp_dim=256,
                         ln_eps=1e-12, lr=1e-3, epoch=1, mini_batch=4),
}


def _load(golden_dir, which):
    return np.load(os.path.join(golden_dir, f"golden_{which}.npz"))


@pytest.mark.parametrize("which", ["lucid", "cls", "hg"])
def test_forward_matches_reference(golden_dir, which):
    G, cfg = _load(golden_dir, which), CFGS[which]
    P = O.init_params(cfg, seed=7)
    rng = np.random.default_rng(11)
    state = np.float32(rng.integers(0, 256, (16, 4, 84, 84), dtype=np.uint8)) / 255.0
    with torch.no_grad():
        pol, ve, vi = O.actor_critic_forward(P, torch.tensor(state), cfg)
    np.testing.assert_allclose(pol.numpy(), G["fwd_policy"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(ve.numpy(), G["fwd_value_ext"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(vi.numpy(), G["fwd_value_int"], rtol=1e-4, atol=1e-6)
    if which == "lucid":
        with torch.no_grad():
            fe = O.lucid_vit(P, torch.tensor(state), O.EXPLORATIVE, cfg).numpy()
            fx = O.lucid_vit(P, torch.tensor(state), O.EXPLOITATIVE, cfg).numpy()
        np.testing.assert_allclose(fe, G["fwd_feat_explorative"], rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(fx, G["fwd_feat_exploitative"], rtol=1e-4, atol=1e-5)
    # get_action: bit-exact action indices for the same uniform draw (agents.py:206-208)
    np.random.seed(5)
    a, v1, v2, lg = O.get_action(P, state, cfg)
    assert np.array_equal(a, G["act_action"])
    np.testing.assert_allclose(lg, G["act_logits"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(v1, G["act_value_ext"], rtol=1e-4, atol=1e-6)
    obs = rng.normal(0, 1, (5, 1, 84, 84)).clip(-5, 5)
    np.testing.assert_allclose(O.intrinsic_reward(P, obs), G["intrinsic_reward"], rtol=1e-4)


def test_numerics_match_reference_bit_exact(golden_dir):
    """GAE, reward filter / rms, obs rms, normalisation: float64 numpy, must be bit-identical."""
    G = _load(golden_dir, "lucid")
    roll = O.synth_rollout(E=6, T=16, seed=3)
    st, rw, ac, dn, no, vex, vin, po = O.relayout_rollout(
        16, 6, roll["total_state"], roll["total_reward"], roll["total_action"], roll["total_done"],
        roll["total_next_obs"], roll["total_ext_values"], roll["total_int_values"], roll["total_policy"])
    et, ea = O.make_train_data(rw, dn, vex, 0.999, 16, 6)
    assert np.array_equal(et, G["gae_ext_target"]) and np.array_equal(ea, G["gae_ext_adv"])
    ir = roll["total_int_reward"].reshape([16, 6]).transpose().reshape([6, 16])
    rrm, flt = O.RunningMeanStd(), O.RewardForwardFilter(0.99)
    O.normalize_int_reward(ir, flt, rrm)
    irn = O.normalize_int_reward(ir * 2, flt, rrm)
    # second call normalises ir*2; the golden normalised ir with the same final var
    irn = ir.copy()
    irn /= np.sqrt(rrm.var)
    assert np.array_equal(irn, G["int_reward_norm"])
    assert np.array_equal(np.array([rrm.mean, rrm.var, rrm.count]), G["reward_rms"])
    assert np.array_equal(flt.rewems, G["filt_rewems"])
    it, ia = O.make_train_data(irn, np.zeros_like(irn), vin, 0.99, 16, 6)
    assert np.array_equal(it, G["gae_int_target"]) and np.array_equal(ia, G["gae_int_adv"])
    orm = O.RunningMeanStd(shape=(1, 1, 84, 84))
    orm.update(no)
    orm.update(no[::2] * 0.5 + 3.0)
    assert np.array_equal(orm.mean, G["obs_rms_mean"]) and np.array_equal(orm.var, G["obs_rms_var"])
    assert orm.count == float(G["obs_rms_count"])
    assert np.array_equal(O.normalize_obs(no[:3], orm), G["obs_norm"])


@pytest.mark.parametrize("which", ["lucid", "hg"])
def test_train_model_matches_reference(golden_dir, which):
    """Full update (permutation, RND mask, loss, backward, Adam) vs RNDAgent.train_model."""
    G, cfg = _load(golden_dir, which), CFGS[which]
    P = O.init_params(cfg, seed=7)
    E, T = 2, 16
    roll = O.synth_rollout(E=E, T=T, seed=21)
    orm, rrm, flt = O.RunningMeanStd(shape=(1, 1, 84, 84)), O.RunningMeanStd(), O.RewardForwardFilter(cfg.int_gamma)
    args = O.prepare_update(cfg, T, E, roll, orm, rrm, flt)
    np.random.seed(123)
    torch.manual_seed(123)
    log = O.train_model(P, cfg, *args)
    assert len(log) == int(G["upd_n_steps"])
    worst = 0.0
    for k, v in P.items():
        if k.endswith("attention.key.bias"):
            # softmax is invariant to a key bias: its true gradient is exactly 0, the computed one is
            # rounding noise, and Adam turns noise into +-lr steps -- not a parity signal.
            continue
        d, g = digest(v.detach().numpy()), G["upd/" + k]
        err = np.abs(d - g).max() / (np.abs(g).max() + 1e-12)
        worst = max(worst, err)
        assert err < 2e-4, (k, err)
    # the update must actually have moved the weights (lr 1e-3, Adam) -- guards a vacuous pass
    P0 = O.init_params(cfg, seed=7)
    k = "model.actor.2.weight"
    assert np.abs(digest(P0[k].numpy()) - G["upd/" + k]).max() > 1e-4


def hg_full_cfg():
    return O.OracleConfig(impl="hg", patch=12, dim=1024, depth=12, heads=16, dim_head=64, mlp_dim=3072, ln_eps=1e-12,
                          lr=1e-4, epoch=1, mini_batch=4)


def hg_full_inputs():
    rng = np.random.default_rng(31)
    state = np.float32(rng.integers(0, 256, (8, 4, 84, 84), dtype=np.uint8)) / 255.0
    w = rng.normal(size=(8, 18)).astype(np.float32)
    return state, w


def test_hg_shipped_size_forward_and_gradients_match_reference(golden_dir):
    """The HF-style ViT at its SHIPPED size (1024 / 12 layers / 16 heads / 3072, vit_hg.py:277-374, model.py:200-220):
    outputs and every parameter gradient of the oracle vs the unmodified reference (`make_golden.py hg_full`)."""
    G, cfg = _load(golden_dir, "hg_full"), hg_full_cfg()
    P = O.init_params(cfg, seed=7)
    state, w = hg_full_inputs()
    for k in P:
        if k.startswith("model."):
            P[k].requires_grad_(True)
    torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))
    pol, ve, vi = O.actor_critic_forward(P, torch.tensor(state), cfg)
    np.testing.assert_allclose(pol.detach().numpy(), G["fwd_policy"], rtol=2e-4, atol=1e-6)
    np.testing.assert_allclose(ve.detach().numpy(), G["fwd_value_ext"], rtol=2e-4, atol=1e-6)
    np.testing.assert_allclose(vi.detach().numpy(), G["fwd_value_int"], rtol=2e-4, atol=1e-6)
    ((pol * torch.tensor(w)).sum() + 3.0 * ve.sum() + 2.0 * vi.sum()).backward()
    keys = [k[len("grad/"):] for k in G.files if k.startswith("grad/")]
    assert len(keys) > 190
    gmax = max(float(G["grad/" + k][0]) for k in keys)
    for k in keys:
        assert P[k].grad is not None, k
        d, g = digest(P[k].grad.numpy()), G["grad/" + k]
        if k.endswith("attention.key.bias"):
            continue                                  # exactly-zero true gradient (softmax shift invariance): rounding noise
        assert np.abs(d - g).max() <= 5e-4 * np.abs(g).max() + 1e-7 * gmax, (k, np.abs(d - g).max(), np.abs(g).max())
    for k in P:                                       # tensors the reference leaves without a gradient stay without one
        if k.startswith("model.") and ("grad/" + k) not in G.files:
            assert P[k].grad is None or float(P[k].grad.abs().max()) == 0.0, k
