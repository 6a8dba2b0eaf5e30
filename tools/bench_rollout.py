"""Rollout-side latency: RNDAgent.get_action (agents.py:187-195) and compute_intrinsic_reward
(agents.py:210-218) at E envs, numpy in -> numpy out, as train.py:605/:665/:702 call them once per env step.

    python tools/bench_rollout.py [E] [dropout]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import eavit_b200  # noqa: F401
from eavit_b200 import _lib, agents, config, utils

E = int(sys.argv[1]) if len(sys.argv) > 1 else 128
drop = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
T, A = 128, 18
config.load_config(os.path.join(ROOT, "configs", "expGlados3_lucidrains_explorative.conf"), ViTlucidrains_dropout=drop,
                   ViTlucidrains_emb_dropout=drop)
c = config.default_config
utils.set_seed(42)
agent = agents.RNDAgent(84, A, utils.Env_action_space_type.DISCRETE, E, T, float(c["Gamma"]), GAE_Lambda=float(c["GAELambda"]),
                        learning_rate=float(c["LearningRate"]), ent_coef=float(c["Entropy"]), epoch=4, batch_size=E * T // 32,
                        ppo_eps=float(c["PPOEps"]), use_cuda=True, representation_lr_method="None", device="cuda:0",
                        logger=utils.Logger())
rng = np.random.default_rng(0)
states = (rng.integers(0, 256, (E, 4, 84, 84), dtype=np.uint8) / np.float32(255.0)).astype(np.float32)
obs = rng.normal(0, 1, (E, 1, 84, 84)).clip(-5, 5)          # float64, as train.py:666 produces it


def timed(name, fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    print(f"{name:46s} {dt * 1e3:8.3f} ms/call   {(_lib.launch_count() - l0) / n:6.1f} kernel launches/call   "
          f"{E / dt:10.0f} env-steps/s", flush=True)
    return dt


print(f"E = {E}, dropout = {drop}, mode = train (fact 6: the reference rolls out in train mode)")
a = timed("get_action(np.float32[E,4,84,84])", lambda: agent.get_action(states))
b = timed("compute_intrinsic_reward(np.float64[E,1,84,84])", lambda: agent.compute_intrinsic_reward(obs))
sd = torch.from_numpy(states).cuda()
od = torch.from_numpy(obs).cuda()
u8 = torch.randint(0, 256, (E, 4, 84, 84), dtype=torch.uint8)
timed("  get_action(device-resident f32 tensor)", lambda: agent.get_action(sd))
timed("  get_action(host uint8 frames)", lambda: agent.get_action(u8))
timed("  compute_intrinsic_reward(device f64 tensor)", lambda: agent.compute_intrinsic_reward(od))
timed("  H2D only: states f32 pageable -> device", lambda: torch.from_numpy(states).cuda())
rt = agent.runtime()
timed("  rt.sync() only", lambda: rt.sync())
print(f"per env step (1 get_action + 1 intrinsic reward): {(a + b) * 1e3:.3f} ms -> rollout of {T} steps = {(a + b) * T * 1e3:.1f} ms")
if hasattr(agent, "rollout_step"):
    timed("rollout_step (CUDA graph, both calls)", lambda: agent.rollout_step(states, obs))

# ---- update-side glue: DeviceRollout.finish() vs the reference's numpy path (oracle restatement of
# train.py:707-779 / :855) on the same rollout
if os.environ.get("EAVIT_BENCH_GLUE", "1") == "1":
    from eavit_b200 import rollout
    from oracle import oracle as O
    cfg = O.OracleConfig()
    roll = O.synth_rollout(E=E, T=T, seed=1)
    d_obs, d_rrm = utils.RunningMeanStd(shape=(1, 1, 84, 84), usage="obs_rms"), utils.RunningMeanStd(usage="reward_rms")
    d_f = utils.RewardForwardFilter(cfg.int_gamma)
    buf = rollout.DeviceRollout(E, T, A)
    ve, vi = roll["total_ext_values"].reshape(T + 1, E), roll["total_int_values"].reshape(T + 1, E)
    u8s, u8o = roll["total_state"].astype(np.uint8), roll["total_next_obs"].astype(np.uint8)
    t0 = time.perf_counter()
    for t in range(T):
        sl = slice(t * E, (t + 1) * E)
        buf.add(t, u8s[sl], u8o[sl], roll["total_reward"][sl], roll["total_done"][sl], roll["total_action"][sl], ve[t], vi[t],
                roll["total_policy"][sl], roll["total_int_reward"][sl])
    buf.add_last_values(ve[T], vi[T])
    torch.cuda.synchronize()
    t_add = time.perf_counter() - t0
    for _ in range(2):
        got = buf.finish(d_obs, d_rrm, d_f, cfg.gamma, cfg.int_gamma, cfg.lam, cfg.ext_coef, cfg.int_coef)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    got = buf.finish(d_obs, d_rrm, d_f, cfg.gamma, cfg.int_gamma, cfg.lam, cfg.ext_coef, cfg.int_coef)
    torch.cuda.synchronize()
    t_fin = time.perf_counter() - t0
    t0 = time.perf_counter()
    ref = O.prepare_update(cfg, T, E, roll, O.RunningMeanStd(shape=(1, 1, 84, 84)), O.RunningMeanStd(), O.RewardForwardFilter(cfg.int_gamma))
    t_ref = time.perf_counter() - t0
    print(f"update glue at {E} envs x {T} steps: {T} x add() (uint8 frames from host) {t_add * 1e3:.1f} ms total, "
          f"finish() {t_fin * 1e3:.2f} ms on device; reference numpy path (oracle port, host) {t_ref * 1e3:.0f} ms")
