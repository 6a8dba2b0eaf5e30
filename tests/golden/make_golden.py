"""Generate golden fixtures by running the UNMODIFIED reference (/root/reference) on CPU.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Writes ``tests/golden/*.npz``.  Inputs and weights come from ``oracle.oracle`` generators
(numpy PCG64, independent of torch's RNG stream), so tests can rebuild them from seeds; only the
reference's OUTPUTS are stored.  Large tensors (gradients / updated weights) are stored as digests
(l2 norm, sum, 64 strided samples) to keep fixtures small.

Import recipe = SURVEY.md section 8(c): argv/env pre-set, matplotlib stubbed, dropout keys = 0.
"""
import configparser
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def digest(a) -> np.ndarray:
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    idx = np.linspace(0, a.size - 1, 64).astype(np.int64)
    return np.concatenate([[np.sqrt((a * a).sum()), a.sum()], a[idx]])


def write_conf(overrides: dict) -> str:
    cp = configparser.ConfigParser()
    cp.optionxform = str
    cp.read(os.path.join(REF, "configs", "demo_config.conf"))
    for k, v in overrides.items():
        cp["DEFAULT"][k] = str(v)
    fd, path = tempfile.mkstemp(suffix=".conf")
    with os.fdopen(fd, "w") as f:
        cp.write(f)
    return path


def import_reference(conf_path: str):
    sys.path.insert(0, REF)
    sys.argv = ["x", "--train", "--config_path", conf_path, "--log_name", "golden"]
    os.environ.update(RANK="0", WORLD_SIZE="1", LOCAL_RANK="0", LOCAL_WORLD_SIZE="1", WANDB_MODE="disabled")
    m, mp = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
    m.pyplot = mp
    sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = m, mp
    os.chdir(tempfile.mkdtemp())
    import agents  # noqa
    import model  # noqa
    import utils  # noqa
    import vit  # noqa
    return agents, model, utils, vit


def HG_FULL_CFG():
    from oracle import oracle as O
    return O.OracleConfig(impl="hg", patch=12, dim=1024, depth=12, heads=16, dim_head=64, mlp_dim=3072, ln_eps=1e-12,
                          lr=1e-4, epoch=1, mini_batch=4)


def hg_full_inputs():
    """8 synthetic 4x84x84 frames (/255) and the weights of the scalar the gradient is taken of."""
    rng = np.random.default_rng(31)
    state = np.float32(rng.integers(0, 256, (8, 4, 84, 84), dtype=np.uint8)) / 255.0
    w = rng.normal(size=(8, 18)).astype(np.float32)
    return state, w


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "lucid"
    import torch
    from oracle import oracle as O

    torch.set_num_threads(8)
    if which == "lucid":
        conf = write_conf({"ViTlucidrains_dropout": 0.0, "ViTlucidrains_emb_dropout": 0.0, "NumStep": 16,
                           "MiniBatch": 4, "Epoch": 2, "LearningRate": 0.001})
        cfg = O.OracleConfig(lr=1e-3, epoch=2, mini_batch=4)
    elif which == "cls":
        conf = write_conf({"ViTlucidrains_dropout": 0.0, "ViTlucidrains_emb_dropout": 0.0,
                           "ViTlucidrains_use_explorativeAttn": "False"})
        cfg = O.OracleConfig(use_explorative=False)
    elif which == "hg":
        conf = write_conf({"ViT_implementation_type": 1, "ViTHG_hidden_size": 128, "ViTHG_num_hidden_layers": 2,
                           "ViTHG_num_attention_heads": 2, "ViTHG_intermediate_size": 256,
                           "ViTHG_PreProcHeight": 84, "ViTHG_StateStackSize": 4,
                           "extracted_feature_embedding_dim": 128, "NumStep": 16, "MiniBatch": 4, "Epoch": 1})
        cfg = O.OracleConfig(impl="hg", patch=12, dim=128, depth=2, heads=2, dim_head=64, mlp_dim=256,
                             ln_eps=1e-12, lr=1e-3, epoch=1, mini_batch=4)
    elif which == "hg_full":
        # the SHIPPED vit_hg size (configs/vit_hg_explorative.conf here; reference keys ViTHG_* of
        # configs/expGlados3/Montezuma/config_originalRND_NoSSL_VitExplorativeAttnLucidrains.conf:36-47): 1024 / 12 L / 16 h / 3072
        conf = write_conf({"ViT_implementation_type": 1, "ViTHG_hidden_size": 1024, "ViTHG_num_hidden_layers": 12,
                           "ViTHG_num_attention_heads": 16, "ViTHG_intermediate_size": 3072,
                           "ViTHG_PreProcHeight": 84, "ViTHG_StateStackSize": 4, "ViTHG_patch_size": 12,
                           "ViTHG_hidden_dropout_prob": 0.0, "ViTHG_attention_probs_dropout_prob": 0.0,
                           "extracted_feature_embedding_dim": 1024, "NumStep": 16, "MiniBatch": 4, "Epoch": 1})
        cfg = HG_FULL_CFG()
    elif which == "init":
        conf = write_conf({"ViTlucidrains_dropout": 0.0, "ViTlucidrains_emb_dropout": 0.0})
        cfg = O.OracleConfig()
    else:
        raise SystemExit(which)
    out_dir = HERE
    agents, model, utils, vit = import_reference(conf)
    from utils import Logger, Env_action_space_type

    if which in ("hg", "hg_full"):
        import vit_hg
        # transformers 5.x dropped get_head_mask (SURVEY 8c shim); arithmetic unchanged
        vit_hg.ViT_ExplorativeAttn.get_head_mask = lambda self, hm, n, *a, **k: [None] * n

    if which == "init":
        # initial weights of the reference constructors under set_seed(42) (utils.py:173), as digests
        utils.set_seed(42)
        logger = Logger(file_log_path="./logs/golden", tb_log_path="./logs/tb")
        agent = agents.RNDAgent(84, 18, Env_action_space_type.DISCRETE, 2, 16, 0.999, use_cuda=False, use_noisy_net=False,
                                representation_lr_method="None", device="cpu", logger=logger)
        np.savez_compressed(os.path.join(out_dir, "golden_init.npz"), **{k: digest(v.numpy()) for k, v in agent.state_dict().items()})
        print("wrote golden_init.npz")
        return
    E, T = 2, 16
    N = E * T
    logger = Logger(file_log_path="./logs/golden", tb_log_path="./logs/tb")
    agent = agents.RNDAgent(84, cfg.n_actions, Env_action_space_type.DISCRETE, E, T, cfg.gamma, GAE_Lambda=cfg.lam,
                            learning_rate=cfg.lr, ent_coef=cfg.ent_coef, max_grad_norm=0.5, epoch=cfg.epoch,
                            batch_size=N // cfg.mini_batch, ppo_eps=cfg.ppo_eps, use_cuda=False,
                            use_noisy_net=False, representation_lr_method="None", device="cpu", logger=logger)
    P = O.init_params(cfg, seed=7)
    sd = agent.state_dict()
    assert set(sd.keys()) == set(P.keys()), (set(sd.keys()) ^ set(P.keys()))
    for k in sd:
        assert tuple(sd[k].shape) == tuple(P[k].shape), k
    agent.load_state_dict({k: v.clone() for k, v in P.items()}, strict=True)
    agent.set_mode("train")   # as train.py:272; dropout keys are 0 so train == eval numerics

    if which == "hg_full":
        # forward + full backward of the shipped-size HF-style ViT on 8 samples: outputs and per-tensor gradient digests
        state, w = hg_full_inputs()
        agent.optimizer.zero_grad()
        pol, ve, vi = agent.model(torch.tensor(state))
        ((pol * torch.tensor(w)).sum() + 3.0 * ve.sum() + 2.0 * vi.sum()).backward()
        G = {"fwd_policy": pol.detach().numpy(), "fwd_value_ext": ve.detach().numpy(), "fwd_value_int": vi.detach().numpy()}
        for k, p_ in agent.named_parameters():
            if k.startswith("model.") and p_.grad is not None:
                G["grad/" + k] = digest(p_.grad.numpy())
        np.savez_compressed(os.path.join(out_dir, "golden_hg_full.npz"), **G)
        print("wrote golden_hg_full.npz", len(G), "entries")
        return

    G = {}
    # ---- forward: CnnActorCriticNetwork / get_action / compute_intrinsic_reward -----------------
    rng = np.random.default_rng(11)
    state_u8 = rng.integers(0, 256, (16, 4, 84, 84), dtype=np.uint8)
    state = np.float32(state_u8) / 255.0
    with torch.no_grad():
        pol, ve, vi = agent.model(torch.tensor(state))
    G["fwd_policy"], G["fwd_value_ext"], G["fwd_value_int"] = pol.numpy(), ve.numpy(), vi.numpy()
    if which in ("lucid",):
        with torch.no_grad():
            G["fwd_feat_explorative"] = agent.model.feature(torch.tensor(state), attn_type=vit.ViT_Attn.EXPLORATIVE_ATTN).numpy()
            G["fwd_feat_exploitative"] = agent.model.feature(torch.tensor(state), attn_type=vit.ViT_Attn.EXPLOITATIVE_ATTN).numpy()
    np.random.seed(5)
    with torch.no_grad():
        a, v1, v2, lg = agent.get_action(state)
    G["act_action"], G["act_value_ext"], G["act_value_int"], G["act_logits"] = a, v1, v2, lg
    obs = rng.normal(0, 1, (5, 1, 84, 84)).clip(-5, 5)
    with torch.no_grad():
        G["intrinsic_reward"] = agent.compute_intrinsic_reward(obs)

    # ---- numerics: GAE / RMS / filter --------------------------------------------------------------
    if which == "lucid":
        roll = O.synth_rollout(E=6, T=16, seed=3)
        st, rw, ac, dn, no, vex, vin, po = O.relayout_rollout(
            16, 6, roll["total_state"], roll["total_reward"], roll["total_action"], roll["total_done"],
            roll["total_next_obs"], roll["total_ext_values"], roll["total_int_values"], roll["total_policy"])
        et, ea = utils.make_train_data(rw, dn, vex, 0.999, 16, 6)
        ir = roll["total_int_reward"].reshape([16, 6]).transpose().reshape([6, 16])
        reward_rms = utils.RunningMeanStd(usage="reward_rms")
        filt = utils.RewardForwardFilter(0.99)
        for rep in range(2):   # filter/rms state carries across updates
            per_env = np.array([filt.update(r) for r in (ir * (rep + 1)).T])
            reward_rms.update_from_moments(np.mean(per_env), np.std(per_env) ** 2, len(per_env))
        irn = ir.copy()
        irn /= np.sqrt(reward_rms.var)
        it, ia = utils.make_train_data(irn, np.zeros_like(irn), vin, 0.99, 16, 6)
        obs_rms = utils.RunningMeanStd(shape=(1, 1, 84, 84), usage="obs_rms")
        obs_rms.update(no)
        obs_rms.update(no[::2] * 0.5 + 3.0)
        G.update(gae_ext_target=et, gae_ext_adv=ea, gae_int_target=it, gae_int_adv=ia, int_reward_norm=irn,
                 reward_rms=np.array([reward_rms.mean, reward_rms.var, reward_rms.count]),
                 filt_rewems=filt.rewems, obs_rms_mean=obs_rms.mean, obs_rms_var=obs_rms.var,
                 obs_rms_count=np.array(obs_rms.count),
                 obs_norm=((no[:3] - obs_rms.mean) / np.sqrt(obs_rms.var)).clip(-5, 5))

    # ---- update: RNDAgent.train_model ---------------------------------------------------------------
    if which in ("lucid", "hg"):
        roll = O.synth_rollout(E=E, T=T, seed=21)
        orm, rrm, flt = O.RunningMeanStd(shape=(1, 1, 84, 84)), O.RunningMeanStd(), O.RewardForwardFilter(cfg.int_gamma)
        args = O.prepare_update(cfg, T, E, roll, orm, rrm, flt)
        utils.set_seed(123)
        agent.train_model(*args, 1)
        new_sd = agent.state_dict()
        for k in new_sd:
            G["upd/" + k] = digest(new_sd[k].numpy())
        G["upd_n_steps"] = np.array(cfg.epoch * cfg.mini_batch)

    np.savez_compressed(os.path.join(out_dir, f"golden_{which}.npz"), **G)
    print("wrote", f"golden_{which}.npz", {k: np.asarray(v).shape for k, v in list(G.items())[:12]})


if __name__ == "__main__":
    main()
