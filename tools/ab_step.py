#!/usr/bin/env python
"""Same-process A/B of the cfg3 optimiser step (captured graph, rollout resident) over engine switches that are plain
attributes of the encoder, e.g.  python tools/ab_step.py fuse_embed_bwd=1 fuse_embed_bwd=0 [--steps 100 --rounds 3].
Configurations are interleaved over the rounds (clock drift shows as spread inside one configuration); the captured step
graph is dropped between configurations.  CUDA events around `steps` replays."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="+", help="attr=value[,attr=value] per configuration (encoder attributes, ints; env:NAME=value for environment switches)")
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--rounds", type=int, default=3)
    a = ap.parse_args()
    torch.cuda.set_device(0)
    E, T, A = bench.E_PER_GPU, bench.T, bench.A
    agent = bench.make_agent(E)
    rt = agent.runtime()
    dev = rt.device
    N, B = E * T, agent.batch_size
    g = torch.Generator(device=dev).manual_seed(1)
    R = dict(states=torch.randint(0, 256, (N, 4, 84, 84), dtype=torch.uint8, device=dev, generator=g),
             te=torch.randn(N, device=dev, generator=g), ti=torch.randn(N, device=dev, generator=g),
             adv=torch.randn(N, device=dev, generator=g), y=torch.randint(0, A, (N,), device=dev, generator=g),
             obs=torch.randn(N, 1, 84, 84, device=dev, generator=g).clamp_(-5, 5), old=torch.randn(N, A, device=dev, generator=g))
    perm = torch.randperm(N, device=dev, generator=g)
    mask = (torch.rand(B, device=dev, generator=g) < 0.25).float()
    n_mb = N // B
    rt.sync()
    res = {c: [] for c in a.configs}
    for rnd in range(a.rounds):
        for c in a.configs:
            for kv in c.split(","):
                k, v = kv.split("=")
                if k.startswith("env:"):                      # an environment switch read when the step is captured
                    os.environ[k[4:]] = v
                    continue
                assert hasattr(rt.encoder, k), k
                setattr(rt.encoder, k, type(getattr(rt.encoder, k))(int(v)))
            agent.__dict__.pop("_step_graphs", None)
            for i in range(5):
                agent.train_step(R, perm[B * (i % n_mb): B * (i % n_mb + 1)], mask)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(a.steps):
                agent.train_step(R, perm[B * (i % n_mb): B * (i % n_mb + 1)], mask)
            e1.record()
            torch.cuda.synchronize()
            res[c].append(e0.elapsed_time(e1) / a.steps)
            print(f"round {rnd} {c:30s} {res[c][-1]:.4f} ms/step", flush=True)
    print(json.dumps({k: {"ms": v, "min": min(v)} for k, v in res.items()}))


if __name__ == "__main__":
    main()
