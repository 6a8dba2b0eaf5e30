"""The optimiser step as one captured CUDA graph (RNDAgent._train_step_graphed) and the static rollout buffers behind it."""
import numpy as np
import pytest
import torch

from oracle import oracle as O
from test_gpu_model import CFGS, make_agent, _batch, rel

pytestmark = pytest.mark.gpu


def _run(monkeypatch, graph: str, steps=3, **conf):
    monkeypatch.setenv("EAVIT_STEP_GRAPH", graph)
    cfg = CFGS["lucid"]
    agent, P = make_agent(cfg, 2, 16, **conf)
    args = _batch(cfg, 2, 16)
    R = agent.upload_rollout(*args)
    # a small step size keeps the two trajectories comparable: Adam turns gradients that are rounding noise into +-lr steps
    # whose sign depends on the summation order of the split-K red.adds, and at lr = 1e-3 that chaos reaches the third
    # step's gradient at the percent level
    agent.optimizer.param_groups[0]["lr"] = 1e-6
    stats = torch.zeros(steps, 16, device="cuda")
    for k in range(steps):
        idx = (torch.arange(8, device="cuda") * 3 + k) % 32
        mask = torch.tensor(((np.arange(8) + k) % 2).astype(np.float32)).cuda()
        agent.train_step(R, idx, mask, stats[k])
    torch.cuda.synchronize()
    return agent, stats.cpu().numpy()


def test_graph_replay_equals_eager_steps(monkeypatch):
    """Three optimiser steps replayed from the captured graph == the same three steps launched eagerly (weights, Adam
    moments, per-step loss terms), up to the summation order of the split-K weight gradients."""
    from eavit_b200 import _lib
    n0 = _lib.launch_count()
    a_g, s_g = _run(monkeypatch, "1")
    n_graph = _lib.launch_count() - n0
    assert "_step_graphs" in a_g.__dict__ and len(a_g._step_graphs) == 1
    n0 = _lib.launch_count()
    a_e, s_e = _run(monkeypatch, "0")
    n_eager = _lib.launch_count() - n0
    assert "_step_graphs" not in a_e.__dict__
    np.testing.assert_allclose(s_g, s_e, rtol=2e-3, atol=1e-5)            # per-step loss terms (bf16 activations downstream of
                                                                          # order-dependent fp32 red.adds: not bitwise run to run)
    st_g, st_e = a_g.runtime().store, a_e.runtime().store
    assert int(st_g.step.item()) == int(st_e.step.item()) == 3
    assert rel(st_g.grad.cpu().numpy(), st_e.grad.cpu().numpy()) < 5e-3   # the last step's gradient
    assert rel(st_g.m.cpu().numpy(), st_e.m.cpu().numpy()) < 5e-3         # Adam moments carried through the replays
    assert (st_g.flat - st_e.flat).abs().max().item() <= 2.1e-6 * 3       # at most lr per step and element
    P0 = O.init_params(CFGS["lucid"], seed=7)
    k0 = "model.feature.transformer.layers.0.0.to_qkv.weight"
    assert (st_g.w(k0).cpu() - P0[k0]).abs().max().item() > 1e-6          # ... and the replays did update the weights
    # launch accounting: replays are counted by their kernel nodes (eager: every launch; graph: warm-up + capture + 3 replays)
    assert n_graph >= n_eager and n_eager > 3 * 100


def test_graph_replays_draw_fresh_dropout_masks(monkeypatch):
    """With dropout active the captured seeds are constants; the graph bumps the device epoch word, so two replays on the
    same minibatch see different masks (different loss terms), while eval mode (no dropout) reproduces itself."""
    monkeypatch.setenv("EAVIT_STEP_GRAPH", "1")
    cfg = CFGS["lucid"]
    agent, P = make_agent(cfg, 2, 16, ViTlucidrains_dropout=0.1, ViTlucidrains_emb_dropout=0.1)
    args = _batch(cfg, 2, 16)
    R = agent.upload_rollout(*args)
    agent.optimizer.param_groups[0]["lr"] = 0.0                  # keep the weights where they are
    idx, mask = torch.arange(8, device="cuda"), torch.ones(8, device="cuda")
    s = torch.zeros(3, 16, device="cuda")
    for k in range(3):
        agent.train_step(R, idx, mask, s[k])
    s = s.cpu().numpy()
    assert abs(s[1, 2] - s[2, 2]) > 1e-6 * abs(s[1, 2])          # critic loss differs between replays: fresh masks
    agent.set_mode("eval")
    t = torch.zeros(2, 16, device="cuda")
    for k in range(2):
        agent.train_step(R, idx, mask, t[k])
    t = t.cpu().numpy()
    np.testing.assert_allclose(t[0, 1:6], t[1, 1:6], rtol=1e-5)


def test_upload_rollout_keeps_addresses_and_never_adopts_caller_tensors():
    cfg = CFGS["lucid"]
    agent, P = make_agent(cfg, 2, 16)
    args = _batch(cfg, 2, 16)
    R1 = agent.upload_rollout(*args)
    ptrs = {k: v.data_ptr() for k, v in R1.items()}
    R2 = agent.upload_rollout(*args)
    assert {k: v.data_ptr() for k, v in R2.items()} == ptrs      # same buffers from update to update: the step graph stays valid
    # CUDA tensors passed by the caller are copied, never adopted: a later upload must not write into the caller's memory
    dev_args = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in args]
    dev_args[0] = (dev_args[0] * 255).round().to(torch.uint8)     # raw frames: a new (name, shape, dtype) signature
    mine = dev_args[0].clone()
    R3 = agent.upload_rollout(*dev_args)
    assert R3["states"].data_ptr() != dev_args[0].data_ptr()
    other = [a.clone() for a in dev_args]
    other[0].fill_(7)
    agent.upload_rollout(*other)
    assert torch.equal(dev_args[0], mine)
    assert int(R3["states"][0, 0, 0, 0]) == 7                     # ... but the agent's own buffer was refreshed in place
