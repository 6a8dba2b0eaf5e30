"""cfg4 sanity at full size: vit_hg (hidden 1024, 12 layers, 16 heads x 64, mlp 3072, patch 12 -> 50 tokens), 128 envs x 128
steps per GPU, minibatch 512: a few optimiser steps, ms/step and the per-kernel table."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import eavit_b200  # noqa
from eavit_b200 import agents, config, ops, utils

E, T, A = 128, 128, 18
config.load_config(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "configs", "vit_hg_explorative.conf"))
c = config.default_config
N = E * T
B = N // int(c["MiniBatch"])
utils.set_seed(42)
agent = agents.RNDAgent(84, A, utils.Env_action_space_type.DISCRETE, E, T, float(c["Gamma"]), GAE_Lambda=float(c["GAELambda"]),
                        learning_rate=float(c["LearningRate"]), ent_coef=float(c["Entropy"]), epoch=1, batch_size=B,
                        ppo_eps=float(c["PPOEps"]), use_cuda=True, representation_lr_method="None", device="cuda:0", logger=utils.Logger())
dev = "cuda"
R = dict(states=torch.randint(0, 256, (N, 4, 84, 84), dtype=torch.uint8, device=dev),
         te=torch.randn(N, device=dev), ti=torch.randn(N, device=dev), adv=torch.randn(N, device=dev),
         y=torch.randint(0, A, (N,), device=dev), obs=torch.randn(N, 1, 84, 84, device=dev).clamp_(-5, 5),
         old=torch.randn(N, A, device=dev))
perm = torch.randperm(N, device=dev)
mask = (torch.rand(B, device=dev) < 0.25).float()
rt = agent.runtime()
rt.sync()
print("params", rt.store.numel, "B", B, "tokens", 2 * B * 50, flush=True)
for i in range(3):
    agent.train_step(R, perm[B * i: B * (i + 1)], mask)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
n = 6
for i in range(n):
    agent.train_step(R, perm[B * (i % 8): B * (i % 8 + 1)], mask)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"vit_hg cfg4 shape: {ms:.2f} ms/step, {B / ms * 1e3:.0f} samples/s, model {76.43e9 * B / ms / 1e9:.0f} TFLOP/s", flush=True)
ops.profile_start()
agent.train_step(R, perm[:B], mask)
tab = ops.profile_stop()
rows = sorted(((l, n_, t, f) for l, (n_, t, f) in tab.items()), key=lambda r: -r[2])
for l, n_, t, f in rows[:22]:
    print(f"{l[:64]:64s} {n_:4d} {t:8.3f} ms {(f * n_ / t / 1e9) if f else 0:7.0f} TF/s")
print("sum", sum(r[2] for r in rows))
