#!/usr/bin/env python
"""REJECTED EXPERIMENT -- needs tools/rejected/r2_stream_scheduling_switches.patch applied (the switches it toggles are not
in the shipped code; result: profiles/r2_ab_stream_scheduling.json).

A/B of the step's stream scheduling switches on one GPU (cfg3 shapes, rollout resident, captured step graph).

    python tools/ab_sched.py [--steps 100] [--rounds 2]

Each configuration = a set of environment switches read when the step is captured (agents.RNDAgent._train_step_eager):
EAVIT_RND_FORK (where the towers' stream forks off), EAVIT_STREAM_PRIO (capture stream high / side streams low priority),
EAVIT_DW_SIDE (small weight-gradient GEMMs on a third stream).  The captured graph, the side streams and the capture stream
are dropped between configurations so that each one is captured afresh.  CUDA events around `steps` replays, configurations
interleaved over `rounds` passes (clock drift shows up as a spread between the passes of one configuration).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402

CONFIGS = [
    ("base", dict(EAVIT_RND_FORK="start", EAVIT_STREAM_PRIO="0", EAVIT_DW_SIDE="0")),
    ("prio", dict(EAVIT_RND_FORK="start", EAVIT_STREAM_PRIO="1", EAVIT_DW_SIDE="0")),
    ("mid", dict(EAVIT_RND_FORK="mid", EAVIT_STREAM_PRIO="0", EAVIT_DW_SIDE="0")),
    ("mid+prio", dict(EAVIT_RND_FORK="mid", EAVIT_STREAM_PRIO="1", EAVIT_DW_SIDE="0")),
    ("dw", dict(EAVIT_RND_FORK="start", EAVIT_STREAM_PRIO="0", EAVIT_DW_SIDE="1")),
    ("mid+prio+dw", dict(EAVIT_RND_FORK="mid", EAVIT_STREAM_PRIO="1", EAVIT_DW_SIDE="1")),
    ("prio+dw", dict(EAVIT_RND_FORK="start", EAVIT_STREAM_PRIO="1", EAVIT_DW_SIDE="1")),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--rounds", type=int, default=2)
    ap.add_argument("--out", default=None)
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
    res = {name: [] for name, _ in CONFIGS}
    for rnd in range(a.rounds):
        for name, env in CONFIGS:
            os.environ.update(env)
            agent.__dict__.pop("_step_graphs", None)
            agent.__dict__.pop("_capture_stream", None)
            for k in ("_side", "_side_dw"):
                if hasattr(rt, k):
                    delattr(rt, k)
            for i in range(5):
                agent.train_step(R, perm[B * (i % n_mb): B * (i % n_mb + 1)], mask)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(a.steps):
                agent.train_step(R, perm[B * (i % n_mb): B * (i % n_mb + 1)], mask)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.steps
            res[name].append(ms)
            print(f"round {rnd} {name:14s} {ms:.4f} ms/step", flush=True)
    summary = {k: {"ms": v, "min": min(v)} for k, v in res.items()}
    print(json.dumps(summary))
    if a.out:
        json.dump(summary, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
