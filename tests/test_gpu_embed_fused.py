"""The one-kernel patch embedding (csrc/embed_fused.cu: patchify + LayerNorm(144) + Linear + LayerNorm(256) + token / position
assembly, vit.py:109-114 and :141-158) against the four-launch path it replaces and against the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle as O
from test_gpu_model import CFGS, make_agent, rel

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("which,B,dtype", [("lucid", 5, torch.float32), ("lucid", 16, torch.uint8), ("cls", 7, torch.float32),
                                           ("lucid", 130, torch.uint8)])
def test_fused_embedding_equals_the_four_launch_path(which, B, dtype, monkeypatch):
    cfg = CFGS[which]
    agent, P = make_agent(cfg, 2, 8)
    rt = agent.runtime()
    rng = np.random.default_rng(B)
    u8 = torch.tensor(rng.integers(0, 256, (B + 3, 4, 84, 84), dtype=np.uint8)).cuda()
    img = u8 if dtype == torch.uint8 else (u8.float() / 255.0)
    idx = torch.tensor(rng.permutation(B + 3)[:B]).cuda()                 # minibatch gather by index
    out = {}
    for fused in ("1", "0"):
        monkeypatch.setenv("EAVIT_FUSE_EMBED", fused)
        rt.encoder.buf.clear()
        rt.encoder.forward(img, B, idx)
        torch.cuda.synchronize()
        bf = rt.encoder.buf[B]
        out[fused] = {k: bf.t[k].detach().float().cpu().clone() for k in ("x0", "e0", "pln", "pmean", "prstd", "m3", "r3")}
    a, b = out["1"], out["0"]
    assert rel(a["pmean"], b["pmean"]) < 1e-5 and rel(a["prstd"], b["prstd"]) < 1e-5
    assert rel(a["pln"], b["pln"]) < 4e-3                                  # bf16 values; a statistic that differs in the last bit moves some by one ulp
    assert rel(a["e0"], b["e0"]) < 1e-3
    assert rel(a["m3"], b["m3"]) < 1e-3 and rel(a["r3"], b["r3"]) < 1e-3
    assert rel(a["x0"], b["x0"]) < 1e-3
    # against the oracle's embedding (fp32 torch): both sequences of every sample
    x = (u8.float() / 255.0)[idx].cpu()
    with torch.no_grad():
        if which == "lucid":
            xa = O.lucid_embed(P, x, O.EXPLORATIVE, cfg).reshape(-1, cfg.dim)
            xb = O.lucid_embed(P, x, O.EXPLOITATIVE, cfg).reshape(-1, cfg.dim)
            ref = torch.cat((xa, xb))
        else:
            ref = O.lucid_embed(P, x, O.CLS, cfg).reshape(-1, cfg.dim)
    assert a["x0"].shape == ref.shape
    assert rel(a["x0"], ref) < 5e-3


def test_train_step_gradients_with_and_without_fused_embedding(monkeypatch):
    """The backward consumes what the fused kernel saved (pln, patch statistics, e0, m3, r3): same gradients either way."""
    cfg = CFGS["lucid"]
    agent, P = make_agent(cfg, 2, 16)
    roll = O.synth_rollout(E=2, T=16, seed=9)
    args = O.prepare_update(cfg, 16, 2, roll, O.RunningMeanStd(shape=(1, 1, 84, 84)), O.RunningMeanStd(), O.RewardForwardFilter(cfg.int_gamma))
    R = agent.upload_rollout(*args)
    idx = torch.arange(16, device="cuda")
    mask = torch.ones(16, device="cuda")
    st = agent.runtime().store
    g = {}
    for fused in ("1", "0"):
        monkeypatch.setenv("EAVIT_FUSE_EMBED", fused)
        agent.train_step(R, idx, mask, None, apply=False)
        torch.cuda.synchronize()
        g[fused] = st.grad.detach().cpu().clone()
    assert rel(g["1"], g["0"]) < 3e-3
