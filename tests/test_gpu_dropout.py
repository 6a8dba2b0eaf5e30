"""GPU tests of nn.Dropout (vit.py:31,33,45,56,158) as counter-based masks.

torch's Philox stream cannot be matched by another implementation, so dropout parity is proven with EXPLICIT masks:
``eavit_dropout_mask`` materialises exactly the mask the fused kernels evaluate in place, and the torch / oracle reference
is run with those masks.  Statistical tests cover the keep rate, the scale and the independence of sites / calls.
"""
import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from eavit_b200 import ops as _ops
    return _ops


def rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-20)).item()


def test_mask_statistics(ops):
    p, n, c = 0.1, 4096, 1024
    m = ops.dropout_mask(n, c, p, ops.site_seed(1234, 3))
    vals = torch.unique(m)
    scale = 1.0 / (1.0 - round(p * 65536) / 65536)
    assert vals.numel() == 2 and vals[0].item() == 0.0 and abs(vals[1].item() - scale) < 1e-6
    keep = (m > 0).float()
    assert abs(keep.mean().item() - 0.9) < 1.5e-3                     # 4.2M Bernoulli draws: sigma = 1.5e-4
    assert abs(m.mean().item() - 1.0) < 2e-3                          # unbiased: E[mask] = 1
    assert (keep.mean(0) - 0.9).abs().max().item() < 0.03             # every column, every row close to the rate
    assert (keep.mean(1) - 0.9).abs().max().item() < 0.05
    k = keep - keep.mean()
    for shift_r, shift_c in ((0, 1), (1, 0), (0, 2), (1, 1)):         # neighbouring elements are uncorrelated
        a = k[: n - shift_r, : c - shift_c]
        b = k[shift_r:, shift_c:]
        corr = (a * b).mean().item() / k.var().item()
        assert abs(corr) < 5e-3, (shift_r, shift_c, corr)
    m2 = ops.dropout_mask(n, c, p, ops.site_seed(1234, 4))            # another site / call: independent mask
    both = ((m > 0) & (m2 > 0)).float().mean().item()
    assert abs(both - 0.81) < 3e-3
    assert torch.equal(m, ops.dropout_mask(n, c, p, ops.site_seed(1234, 3)))          # pure function of (seed, r, c)
    sub = ops.dropout_mask(7, 10, p, ops.site_seed(1234, 3), row0=100, col0=33)
    assert torch.equal(sub, m[100:107, 33:43])
    assert (ops.dropout_mask(64, 64, 0.0, 5) == 1.0).all()


def test_gemm_epilogue_dropout(ops):
    torch.manual_seed(0)
    M, N, K, p = 700, 1024, 256, 0.1
    seed = ops.site_seed(77, 11)
    A = torch.randn(M, K, device="cuda").bfloat16()
    B = (torch.randn(N, K, device="cuda") / 16).bfloat16()
    bias = torch.randn(N, device="cuda")
    res = torch.randn(M, N, device="cuda")
    mask = ops.dropout_mask(M, N, p, seed)
    pre = A.float() @ B.float().t() + bias
    out = torch.empty(M, N, device="cuda")
    ops.gemm(A, B, bias=bias, residual=res, out_f32=out, drop_p=p, drop_seed=seed)          # Linear -> Dropout -> + residual
    assert rel(out, pre * mask + res) < 1e-5
    h16, pre16 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16), torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(A, B, bias=bias, act=ops.ACT_GELU, out_bf16=h16, out_pre=pre16, drop_p=p, drop_seed=seed)   # GELU -> Dropout
    assert rel(pre16, pre) < 5e-3
    assert rel(h16, torch.nn.functional.gelu(pre) * mask) < 5e-3
    assert ((h16 == 0) == (mask == 0)).float().mean().item() > 0.999
    aux = torch.randn(M, N, device="cuda").bfloat16()
    x = aux.float().requires_grad_(True)
    torch.nn.functional.gelu(x).sum().backward()
    cs = torch.zeros(N, device="cuda")
    ops.gemm(A, B, act=ops.ACT_GELU_BWD, aux=aux, out_bf16=h16, colsum=cs, drop_p=p, drop_seed=seed)
    ref = (A.float() @ B.float().t()) * x.grad * mask
    assert rel(h16, ref) < 5e-3
    assert rel(cs, ref.sum(0)) < 2e-3


@pytest.mark.parametrize("M", [700, 20000])
def test_gemm_tma_epilogues_dropout(ops, M):
    """The TMA epilogues regenerate the same (row, column) masks in the accumulator's native layout: GELU storing gelu'
    (mask on the activation only), multiply-by-aux + column sums, residual and residual + LayerNorm."""
    torch.manual_seed(M)
    N, K, p = 1024, 256, 0.1
    seed = ops.site_seed(91, 5)
    A = torch.randn(M, K, device="cuda").bfloat16()
    B = (torch.randn(N, K, device="cuda") / 16).bfloat16()
    bias = torch.randn(N, device="cuda")
    mask = ops.dropout_mask(M, N, p, seed)
    pre = (A.float() @ B.float().t() + bias).requires_grad_(True)
    hr = torch.nn.functional.gelu(pre)
    hr.sum().backward()
    h16, g16 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16), torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(A, B, bias=bias, act=ops.ACT_GELU_SAVE_GRAD, out_bf16=h16, out_pre=g16, drop_p=p, drop_seed=seed)
    assert rel(h16, hr.detach() * mask) < 5e-3
    assert rel(g16, pre.grad) < 5e-3
    assert ((h16 == 0) == (mask == 0)).float().mean().item() > 0.999
    cs = torch.zeros(N, device="cuda")
    d16 = torch.empty_like(h16)
    ops.gemm(A, B, act=ops.ACT_MUL_AUX, aux=g16, out_bf16=d16, colsum=cs, drop_p=p, drop_seed=seed)
    want = (A.float() @ B.float().t()) * g16.float() * mask
    assert rel(d16, want) < 5e-3
    assert rel(cs, want.sum(0)) < 3e-3
    # N = 256: residual (+ LayerNorm) epilogues, K = 1024 operand
    A2 = h16
    W2 = (torch.randn(256, N, device="cuda") / 32).bfloat16()
    b2 = torch.randn(256, device="cuda")
    res = torch.randn(M, 256, device="cuda")
    m2 = ops.dropout_mask(M, 256, p, seed)
    y = (A2.float() @ W2.float().t() + b2) * m2 + res
    out = torch.empty(M, 256, device="cuda")
    ops.gemm(A2, W2, bias=b2, residual=res, out_f32=out, drop_p=p, drop_seed=seed)
    assert rel(out, y) < 1e-5
    gam, bet = torch.randn(256, device="cuda") * 0.2 + 1, torch.randn(256, device="cuda") * 0.1
    xn = torch.empty(M, 256, device="cuda", dtype=torch.bfloat16)
    mu, rs = torch.empty(M, device="cuda"), torch.empty(M, device="cuda")
    ops.gemm(A2, W2, bias=b2, residual=res, out_f32=out, out_bf16=xn, ln=(gam, bet, mu, rs, 1e-5), drop_p=p, drop_seed=seed)
    assert rel(out, y) < 1e-5
    assert rel(xn, torch.nn.functional.layer_norm(y, (256,), gam, bet, 1e-5)) < 5e-3
    assert rel(mu, y.mean(1)) < 1e-5


def test_layernorm_bwd_masks_linear_output_gradient(ops):
    torch.manual_seed(1)
    T, D, p = 999, 256, 0.1
    seed = ops.site_seed(5, 2)
    x = torch.randn(T, D, device="cuda")
    g, b = torch.randn(D, device="cuda"), torch.randn(D, device="cuda")
    y = torch.empty(T, D, device="cuda", dtype=torch.bfloat16)
    mean, rstd = torch.empty(T, device="cuda"), torch.empty(T, device="cuda")
    ops.call("eavit_layernorm_fwd", x, D, g, b, y, ops.BF16, D, mean, rstd, T, D, 1e-5)
    dy = torch.randn(T, D, device="cuda").bfloat16()
    dres = torch.randn(T, D, device="cuda")
    dx, dx16 = torch.empty(T, D, device="cuda"), torch.empty(T, D, device="cuda", dtype=torch.bfloat16)
    dg, db, dsum = torch.zeros(D, device="cuda"), torch.zeros(D, device="cuda"), torch.zeros(D, device="cuda")
    ops.call("eavit_layernorm_bwd", dy, ops.BF16, D, x, D, mean, rstd, g, dres, D, dx, D, dx16, D, dg, db, dsum, p, seed, T, D)
    xr = x.clone().requires_grad_(True)
    torch.nn.functional.layer_norm(xr, (D,), g, b, 1e-5).backward(dy.float())
    ref = xr.grad + dres
    mask = ops.dropout_mask(T, D, p, seed)
    assert rel(dx, ref) < 1e-5                              # the residual-stream gradient is NOT masked
    assert rel(dx16, ref * mask) < 5e-3                      # the Linear-output gradient is
    assert rel(dsum, (ref * mask).sum(0)) < 1e-4


def ref_attention_drop(qkv, starts, H, Dh, scale, mask_fn):
    q, k, v = qkv.split(H * Dh, dim=1)
    outs = []
    for s0, s1 in zip(starts[:-1], starts[1:]):
        hs = []
        for h in range(H):
            sl = slice(h * Dh, (h + 1) * Dh)
            pr = ((q[s0:s1, sl] @ k[s0:s1, sl].t()) * scale).softmax(-1)
            hs.append((pr * mask_fn(s0, s1, h)) @ v[s0:s1, sl])
        outs.append(torch.cat(hs, dim=1))
    return torch.cat(outs, dim=0)


@pytest.mark.parametrize("lens,H,Dh", [([196, 197, 196, 197], 8, 32), ([50, 50, 50], 2, 64), ([1, 17, 128, 129, 224], 4, 32),
                                        ([50] * 5, 2, 64), ([64] * 7, 4, 32)])      # the last two run packed (block-diagonal mask)
def test_attention_probability_dropout(ops, lens, H, Dh):
    torch.manual_seed(sum(lens))
    p, seed = 0.1, ops.site_seed(31, 9)
    starts = [0]
    for n in lens:
        starts.append(starts[-1] + n)
    T = starts[-1]
    qkv = (torch.randn(T, 3 * H * Dh, device="cuda") * 1.2).bfloat16()
    dout = torch.randn(T, H * Dh, device="cuda").bfloat16()
    ss = torch.tensor(starts, dtype=torch.int32, device="cuda")
    scale = Dh ** -0.5

    def mask_fn(s0, s1, h):                                   # element (token row, h * 256 + key)
        return ops.dropout_mask(s1 - s0, s1 - s0, p, seed, row0=s0, col0=h * 256)
    x = qkv.float().requires_grad_(True)
    ro = ref_attention_drop(x, starts, H, Dh, scale, mask_fn)
    (ro * dout.float()).sum().backward()
    out = torch.empty(T, H * Dh, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(T, H, device="cuda")
    dqkv = torch.full((T, 3 * H * Dh), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.attention_fwd(qkv, ss, len(lens), max(lens), H, Dh, scale, out, lse, drop_p=p, drop_seed=seed)
    ops.attention_bwd(qkv, out, dout, lse, ss, len(lens), max(lens), H, Dh, scale, dqkv, drop_p=p, drop_seed=seed)
    torch.cuda.synchronize()
    assert rel(out, ro) < 8e-3, rel(out, ro)
    assert torch.isfinite(dqkv.float()).all()
    n = H * Dh
    for name, sl in (("dq", slice(0, n)), ("dk", slice(n, 2 * n)), ("dv", slice(2 * n, 3 * n))):
        e = rel(dqkv[:, sl], x.grad[:, sl])
        assert e < 1.2e-2, (name, e)


KIND = {"emb": 0, "attn_p": 1, "attn_out": 2, "act": 3, "ff_out": 4}


@pytest.mark.parametrize("which", ["lucid", "hg"])
def test_train_step_with_dropout_matches_oracle_with_same_masks(which):
    """One PPO+RND minibatch with dropout 0.1 at every site: loss terms and the total gradient match the oracle run with
    the masks the kernels generated (1e-2 norm-wise, the bf16 tolerance of the dropout-free test)."""
    from test_gpu_model import CFGS, make_agent
    from eavit_b200 import ops, _lib
    cfg = CFGS[which]
    E, T, B = 2, 8, 8
    conf = ({"ViTlucidrains_dropout": 0.1, "ViTlucidrains_emb_dropout": 0.1} if which == "lucid"
            else {"ViTHG_hidden_dropout_prob": 0.1, "ViTHG_attention_probs_dropout_prob": 0.1})
    agent, P = make_agent(cfg, E, T, **conf)
    agent.set_mode("train")
    roll = O.synth_rollout(E=E, T=T, seed=21)
    args = O.prepare_update(cfg, T, E, roll, O.RunningMeanStd(shape=(1, 1, 84, 84)), O.RunningMeanStd(), O.RewardForwardFilter(cfg.int_gamma))
    states, te, ti, y, adv, obs, old = args
    idx = np.arange(B)
    mask = (np.arange(B) % 2).astype(np.float32)
    R = agent.upload_rollout(*args)
    stats = torch.zeros(16, device="cuda")
    rt = agent.runtime()
    agent.train_step(R, torch.from_numpy(idx).cuda(), torch.tensor(mask).cuda(), stats, apply=False)
    torch.cuda.synchronize()
    base = rt._drop_seed0 + rt._drop_calls
    assert rt._drop_calls == 1
    np_, S1 = cfg.n_patches, cfg.n_patches + 1
    hcfg = rt.cfg

    def hook(kind, layer, x, pass_id):
        pr = dict(emb=hcfg.emb_dropout, attn_p=hcfg.attn_dropout, attn_out=hcfg.dropout, act=hcfg.act_dropout, ff_out=hcfg.dropout)[kind]
        if pr <= 0:
            return x
        seed = ops.site_seed(base, layer * 8 + KIND[kind])
        if which == "lucid":
            n = np_ if pass_id == O.EXPLORATIVE else S1
            row0 = 0 if pass_id == O.EXPLORATIVE else B * np_
        else:
            n, row0 = S1, pass_id * B * S1
        if layer == cfg.depth - 1 and kind in ("attn_out", "act", "ff_out") and rt.encoder.prune_last:
            # last layer: only token 0 of each sequence is computed (compact rows = sequence index); other tokens are dead
            m = torch.ones(x.shape)
            seq0 = (0 if pass_id in (O.EXPLORATIVE, 0) else B) if which == "lucid" else pass_id * B
            if which == "lucid" and pass_id == O.EXPLOITATIVE:
                seq0 = B
            m[:, 0, :] = ops.dropout_mask(x.shape[0], x.shape[2], pr, seed, row0=seq0).cpu()
            return x * m
        if kind == "attn_p":                                  # x [b, h, n, n]; element (token row, h*256 + key)
            b, h = x.shape[0], x.shape[1]
            m = torch.stack([torch.stack([ops.dropout_mask(n, n, pr, seed, row0=row0 + bi * n, col0=hi * 256) for hi in range(h)])
                             for bi in range(b)])
        else:                                                 # x [b, n, C]; element (token row, column)
            m = ops.dropout_mask(x.shape[0] * n, x.shape[2], pr, seed, row0=row0).reshape(x.shape)
        return x * m.cpu()

    for k in O.trainable_names(P):
        P[k].requires_grad_(True)
    old_flat = torch.tensor(old).permute(1, 0, 2).contiguous().view(-1, cfg.n_actions)
    O.DROPOUT_HOOK = hook
    try:
        loss, terms, _ = O.ppo_rnd_loss(P, cfg, torch.FloatTensor(states)[idx], torch.FloatTensor(te)[idx], torch.FloatTensor(ti)[idx],
                                        torch.LongTensor(y)[idx], torch.FloatTensor(adv)[idx], torch.FloatTensor(obs)[idx],
                                        old_flat[idx], torch.tensor(mask))
        loss.backward()
    finally:
        O.DROPOUT_HOOK = None
    s = stats.cpu().numpy()
    got = dict(actor=s[1], critic_ext=s[2], critic_int=s[3], entropy=s[4], rnd=s[5])
    for k, v in got.items():
        assert abs(v - terms[k]) <= 1e-2 * max(abs(terms[k]), 1e-3), (k, v, terms[k])
    st = rt.store
    names = [k for k in O.trainable_names(P) if P[k].grad is not None]
    ref = torch.cat([P[k].grad.reshape(-1) for k in names])
    mine = torch.cat([st.g(k).cpu().reshape(-1) for k in names])
    err = float((mine - ref).norm() / ref.norm())
    assert err < 1e-2, err
    # and dropout really was active: the dropout-free oracle gives a different loss
    with torch.no_grad():
        l0, _, _ = O.ppo_rnd_loss(P, cfg, torch.FloatTensor(states)[idx], torch.FloatTensor(te)[idx], torch.FloatTensor(ti)[idx],
                                  torch.LongTensor(y)[idx], torch.FloatTensor(adv)[idx], torch.FloatTensor(obs)[idx], old_flat[idx], torch.tensor(mask))
    assert abs(float(l0) - float(loss.detach())) > 1e-4


def test_eval_mode_disables_dropout():
    from test_gpu_model import CFGS, make_agent
    agent, P = make_agent(CFGS["lucid"], 2, 8, ViTlucidrains_dropout=0.1, ViTlucidrains_emb_dropout=0.1)
    x = torch.rand(4, 4, 84, 84, device="cuda")
    agent.set_mode("eval")
    with torch.no_grad():
        a = agent.model(x)[0].clone()
        b = agent.model(x)[0].clone()
    assert torch.equal(a, b)
    agent.set_mode("train")                                   # the rollout of the reference runs in train mode (fact 6)
    with torch.no_grad():
        c = agent.model(x)[0].clone()
        d = agent.model(x)[0].clone()
    assert not torch.equal(c, d) and not torch.equal(a, c)
