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
    # the fused kernel saves the normalised patch WITHOUT LayerNorm(patch_dim)'s affine (engine: pln_is_xhat -- the backward
    # folds the affine into one weight-gradient GEMM); the four-launch path saves gamma * xhat + beta
    g1 = P["model.feature.to_patch_embedding.1.weight"].detach().float().cpu()
    b1 = P["model.feature.to_patch_embedding.1.bias"].detach().float().cpu()
    assert rt.encoder.fuse_embed_bwd
    assert rel(a["pln"] * g1 + b1, b["pln"]) < 6e-3                        # bf16 values on both sides
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
    emb = [f"model.feature.to_patch_embedding.{k}" for k in ("1.weight", "1.bias", "2.weight", "2.bias", "3.weight", "3.bias")] + \
          ["model.feature.pos_embedding", "model.feature.exploration_token"]
    g, ge = {}, {}
    for fused in ("1", "0"):
        monkeypatch.setenv("EAVIT_FUSE_EMBED", fused)
        agent.train_step(R, idx, mask, None, apply=False)
        torch.cuda.synchronize()
        g[fused] = st.grad.detach().cpu().clone()
        ge[fused] = {n: st.g(n).detach().cpu().clone() for n in emb}
    assert rel(g["1"], g["0"]) < 3e-3
    # the embedding's own tensors one by one: fused forward + folded backward (G = dE^T xhat) against patchify / GEMM / LayerNorm
    # launches with the dX GEMM and the second pass over the frames
    # (two bf16-operand evaluations of the same gradient on 16 samples, each within 1e-2 of the oracle -- test_gpu_model.py
    # checks that per tensor -- so twice that between them)
    for n in emb:
        assert rel(ge["1"][n], ge["0"][n]) < 2e-2, n


def test_patch_ln_fold_bwd():
    """vit.py:111-112 backward without an input gradient: dW, dbias, dgamma, dbeta from G = dY^T xhat and s = colsum(dY)
    against torch autograd (float64) of y = (g1 * xhat + b1) W^T + bias."""
    from eavit_b200.ops import call
    gen = torch.Generator(device="cuda").manual_seed(5)
    R_, N, K = 777, 256, 144
    xh = torch.randn(R_, K, device="cuda", generator=gen, dtype=torch.float64)
    dy = torch.randn(R_, N, device="cuda", generator=gen, dtype=torch.float64)
    W = torch.randn(N, K, device="cuda", generator=gen, dtype=torch.float64).requires_grad_(True)
    bias = torch.zeros(N, device="cuda", dtype=torch.float64, requires_grad=True)
    g1 = torch.randn(K, device="cuda", generator=gen, dtype=torch.float64).requires_grad_(True)
    b1 = torch.randn(K, device="cuda", generator=gen, dtype=torch.float64).requires_grad_(True)
    ((xh * g1 + b1) @ W.t() + bias).backward(dy)
    G = (dy.t() @ xh).float().contiguous()
    sv = dy.sum(0).float().contiguous()
    outs = [torch.randn(N, K, device="cuda", generator=gen), torch.randn(N, device="cuda", generator=gen),
            torch.randn(K, device="cuda", generator=gen), torch.randn(K, device="cuda", generator=gen)]
    base = [o.double().clone() for o in outs]
    call("eavit_patch_ln_fold_bwd", G, sv, W.detach().float().contiguous(), g1.detach().float().contiguous(),
         b1.detach().float().contiguous(), outs[0], outs[1], outs[2], outs[3], N, K)
    torch.cuda.synchronize()
    for o, b0, ref in zip(outs, base, (W.grad, bias.grad, g1.grad, b1.grad)):
        assert rel((o.double() - b0).cpu(), ref.cpu()) < 1e-5


@pytest.mark.parametrize("mode,B,np_,D", [(0, 5, 196, 256), (1, 7, 196, 256), (2, 9, 49, 1024), (0, 70, 16, 384), (2, 3, 4, 128),
                                          (0, 512, 196, 256)])
def test_embed_assemble_bwd_single_pass(mode, B, np_, D):
    """vit.py:141-158 backward: patch-token gradient summed over the sequences, positional / token gradients summed over the
    samples -- the one-pass kernel (a CTA per position and sample chunk) against a torch restatement; accumulation into
    non-zero dpos / dtok buffers (the gradients are +=)."""
    from eavit_b200.ops import call
    g = torch.Generator(device="cuda").manual_seed(mode * 100 + B)
    S1 = np_ + 1
    T = B * np_ + B * S1 if mode == 0 else (B * S1 if mode == 1 else 2 * B * S1)
    dx = torch.randn(T, D, device="cuda", generator=g)
    out = torch.empty(B * np_, D, device="cuda")
    out16 = torch.empty(B * np_, D, device="cuda", dtype=torch.bfloat16)
    dpos0 = torch.randn(S1, D, device="cuda", generator=g)
    dta0, dtb0 = torch.randn(D, device="cuda", generator=g), torch.randn(D, device="cuda", generator=g)
    dpos, dta, dtb = dpos0.clone(), dta0.clone(), dtb0.clone()
    call("eavit_embed_assemble_bwd", dx, mode, B, np_, D, out, out16, dpos, dta, dtb if mode == 2 else None)
    torch.cuda.synchronize()
    d = dx.double()
    if mode == 0:
        a, b = d[: B * np_].view(B, np_, D), d[B * np_:].view(B, S1, D)
        ref_g, ref_pos, ref_ta, ref_tb = a + b[:, 1:], b.sum(0), b[:, 0].sum(0), None
    elif mode == 1:
        a = d.view(B, S1, D)
        ref_g, ref_pos, ref_ta, ref_tb = a[:, 1:], a.sum(0), a[:, 0].sum(0), None
    else:
        a, b = d[: B * S1].view(B, S1, D), d[B * S1:].view(B, S1, D)
        ref_g, ref_pos, ref_ta, ref_tb = a[:, 1:] + b[:, 1:], (a + b).sum(0), a[:, 0].sum(0), b[:, 0].sum(0)
    assert torch.equal(out, ref_g.reshape(B * np_, D).float())                   # one fp32 add per element: exact
    assert torch.equal(out16, ref_g.reshape(B * np_, D).float().bfloat16())
    tol = 1e-6 * max(1.0, B ** 0.5) * 8
    assert (dpos.double() - dpos0.double() - ref_pos).abs().max().item() < tol * 4
    assert (dta.double() - dta0.double() - ref_ta).abs().max().item() < tol * 4
    if mode == 2:
        assert (dtb.double() - dtb0.double() - ref_tb).abs().max().item() < tol * 4
    else:
        assert torch.equal(dtb, dtb0)


@pytest.mark.parametrize("l1", [False, True])
@pytest.mark.parametrize("mode,B,np_,p", [(0, 5, 196, 0.0), (1, 7, 196, 0.0), (0, 70, 16, 0.1), (1, 9, 49, 0.25), (0, 512, 196, 0.0),
                                          (0, 512, 196, 0.1)])
def test_embed_assemble_ln_bwd_single_pass(mode, B, np_, p, l1):
    """The same pass continued through the LayerNorm(256) that ends to_patch_embedding (vit.py:113): de, dgamma, dbeta, the
    Linear's bias gradient and the positional / token gradients against torch autograd in float64.  p > 0: dx is read under the
    embedding-dropout mask (vit.py:158) -- the reference multiplies dx by the materialised mask first.  l1: layer 0's
    pre-attention LayerNorm backward (vit.py:47) runs in front of the pass -- the gradient of the embedding output is
    dres + LN1'(dy), formed per row inside the kernel; its dgamma / dbeta are checked too."""
    from eavit_b200 import ops
    from eavit_b200.ops import call
    seed = 0x1234_5678_9ABC_DEF0 + B
    D = 256
    g = torch.Generator(device="cuda").manual_seed(7 + mode * 100 + B)
    S1 = np_ + 1
    T = B * np_ + B * S1 if mode == 0 else B * S1
    rows = B * np_
    dx = torch.randn(T, D, device="cuda", generator=g)
    e0 = torch.randn(rows, D, device="cuda", generator=g) * 1.7 + 0.3
    gamma = torch.randn(D, device="cuda", generator=g)
    mean = e0.double().mean(1)
    rstd = (e0.double().var(1, unbiased=False) + 1e-5).rsqrt()
    de16 = torch.empty(rows, D, device="cuda", dtype=torch.bfloat16)
    dgam, dbet, dbias = (torch.randn(D, device="cuda", generator=g) for _ in range(3))
    dpos = torch.randn(S1, D, device="cuda", generator=g)
    dtok = torch.randn(D, device="cuda", generator=g)
    base = [t.double().clone() for t in (dgam, dbet, dbias, dpos, dtok)]
    l1_args = (None,) * 7
    if l1:
        dy1 = torch.randn(T, D, device="cuda", generator=g).bfloat16()
        x1 = torch.randn(T, D, device="cuda", generator=g) * 0.8 - 0.2
        gam1 = torch.randn(D, device="cuda", generator=g)
        m1 = x1.double().mean(1)
        r1 = (x1.double().var(1, unbiased=False) + 1e-5).rsqrt()
        dg1, db1 = torch.randn(D, device="cuda", generator=g), torch.randn(D, device="cuda", generator=g)
        base1 = [dg1.double().clone(), db1.double().clone()]
        l1_args = (dy1, x1, m1.float(), r1.float(), gam1, dg1, db1)
    call("eavit_embed_assemble_ln_bwd", dx, mode, B, np_, D, e0, mean.float(), rstd.float(), gamma, de16, dgam, dbet, dbias,
         dpos, dtok, None, float(p), seed, *l1_args)
    torch.cuda.synchronize()
    d = dx.double()
    if l1:
        x1d = x1.double().requires_grad_(True)
        g1d = gam1.double().requires_grad_(True)
        b1d = torch.zeros(D, dtype=torch.float64, device="cuda", requires_grad=True)
        torch.nn.functional.layer_norm(x1d, (D,), g1d, b1d, 1e-5).backward(dy1.double())
        d = d + x1d.grad                                                         # dres + LN1'(dy)
        for got, b0, ref in ((dg1, base1[0], g1d.grad), (db1, base1[1], b1d.grad)):
            err = (got.double() - b0 - ref).abs().max().item()
            assert err < 2e-5 * max(1.0, ref.abs().max().item()) * max(1.0, T ** 0.5 / 30), (err, ref.abs().max().item())
    if p > 0:
        m = ops.dropout_mask(T, D, p, seed)
        assert 0.5 * p < float((m == 0).float().mean()) < 1.5 * p
        d = d * m.double()
    if mode == 0:
        a, b = d[: B * np_].view(B, np_, D), d[B * np_:].view(B, S1, D)
        gg, ref_pos, ref_tok = (a + b[:, 1:]).reshape(rows, D), b.sum(0), b[:, 0].sum(0)
    else:
        a = d.view(B, S1, D)
        gg, ref_pos, ref_tok = a[:, 1:].reshape(rows, D), a.sum(0), a[:, 0].sum(0)
    x = e0.double().requires_grad_(True)
    gm = gamma.double().requires_grad_(True)
    bt = torch.zeros(D, dtype=torch.float64, device="cuda", requires_grad=True)
    y = torch.nn.functional.layer_norm(x, (D,), gm, bt, 1e-5)
    y.backward(gg)
    assert rel(de16.double().cpu(), x.grad.cpu()) < 4e-3                                      # bf16 rounding of the output
    for got, b0, ref in ((dgam, base[0], gm.grad), (dbet, base[1], bt.grad), (dbias, base[2], x.grad.sum(0)),
                         (dpos, base[3], ref_pos), (dtok, base[4], ref_tok)):
        err = (got.double() - b0 - ref).abs().max().item()
        assert err < 2e-5 * max(1.0, ref.abs().max().item()) * max(1.0, rows ** 0.5 / 30), (err, ref.abs().max().item())
