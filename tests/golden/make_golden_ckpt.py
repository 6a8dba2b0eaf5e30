"""Checkpoint fixture written by the UNMODIFIED reference classes (/root/reference/utils.py), for the format test.

Run in the build container only:   python tests/golden/make_golden_ckpt.py
Writes tests/golden/golden_ckpt_aux.pt -- the non-tensor half of a train.py:926-959 checkpoint (pickled obs_rms, reward_rms,
discounted_reward + counters; the state_dict half is covered by test_state_dict_names_and_shapes_match_reference) and
tests/golden/golden_ckpt_aux.npz with the same statistics as plain arrays plus the state after ONE more update."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference, write_conf  # noqa: E402


def main():
    import torch
    conf = write_conf({"ViTlucidrains_dropout": 0.0})
    agents, model, utils, vit = import_reference(conf)
    rng = np.random.default_rng(77)
    obs_rms = utils.RunningMeanStd(shape=(1, 1, 84, 84), usage="obs_rms")
    x0 = rng.integers(0, 256, (96, 1, 84, 84)).astype(np.float64)
    obs_rms.update(x0)                                                  # train.py:127-180 initial statistics
    reward_rms = utils.RunningMeanStd(usage="reward_rms")
    filt = utils.RewardForwardFilter(0.99)
    r = rng.random((8, 16)).astype(np.float32)                         # [E, T] intrinsic rewards
    per_step = np.array([filt.update(r[:, t]) for t in range(16)])     # train.py:736-737
    reward_rms.update_from_moments(np.mean(per_step), np.std(per_step) ** 2, len(per_step))
    ckpt = {"obs_rms": obs_rms, "reward_rms": reward_rms, "discounted_reward": filt, "global_update": 3, "global_step": 384,
            "logger.tb_global_steps": {"a": 1}}
    torch.save(ckpt, os.path.join(HERE, "golden_ckpt_aux.pt"))
    # the same objects one update later (what a resumed run must reproduce bit for bit)
    saved = dict(mean=obs_rms.mean.copy(), var=obs_rms.var.copy(), count=obs_rms.count, r_mean=reward_rms.mean, r_var=reward_rms.var,
                 r_count=reward_rms.count, rewems=filt.rewems.copy())
    x1 = rng.integers(0, 256, (64, 1, 84, 84)).astype(np.float64)
    obs_rms.update(x1)
    r2 = rng.random((8, 16)).astype(np.float32)
    per_step = np.array([filt.update(r2[:, t]) for t in range(16)])
    reward_rms.update_from_moments(np.mean(per_step), np.std(per_step) ** 2, len(per_step))
    np.savez_compressed(os.path.join(HERE, "golden_ckpt_aux.npz"), **saved, x1_seed=77, mean1=obs_rms.mean, var1=obs_rms.var,
                        count1=obs_rms.count, r_mean1=reward_rms.mean, r_var1=reward_rms.var, r_count1=reward_rms.count,
                        rewems1=filt.rewems)
    print("wrote golden_ckpt_aux.pt / .npz")


if __name__ == "__main__":
    main()
