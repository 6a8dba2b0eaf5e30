"""fused Adam, global gradient norm and clipping against torch on known fp32 inputs.

Reference behaviour: torch.optim.Adam(lr) at agents.py:129 / :508, utils.global_grad_norm_ (utils.py:141-170),
nn.utils.clip_grad_norm_ (agents.py:496-499), requires_grad = False on the shared backbone (train.py:261-263).
Everything here goes through the C ABI (``ops.call``) or through the agent's optimiser object."""
import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _bf16_rn(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16)


@pytest.mark.parametrize("n,grad_scale", [(4096, 1.0), (1_000_004, 0.25)])
def test_adam_step_matches_torch_adam(n, grad_scale):
    """eavit_adam_step on random fp32 gradients, 3 steps: p, exp_avg, exp_avg_sq and the bf16 shadow vs torch.optim.Adam
    (grad_scale = 1 / world_size is the data-parallel mean folded into the update)."""
    from eavit_b200 import ops
    g = torch.Generator().manual_seed(3)
    p0 = torch.randn(n, generator=g)
    lr, b1, b2, eps = 1e-3, 0.9, 0.999, 1e-8
    ref_p = torch.nn.Parameter(p0.clone().double())          # float64 torch Adam = the exact recurrence
    opt = torch.optim.Adam([ref_p], lr=lr, betas=(b1, b2), eps=eps)
    ref32 = torch.nn.Parameter(p0.clone())                   # float32 torch Adam = the reference's arithmetic
    opt32 = torch.optim.Adam([ref32], lr=lr, betas=(b1, b2), eps=eps)
    p = p0.clone().cuda()
    m, v = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    sh = torch.zeros(n, dtype=torch.bfloat16, device="cuda")
    step = torch.zeros(1, dtype=torch.int64, device="cuda")
    for k in range(3):
        grad = torch.randn(n, generator=g) * (10.0 ** (k - 1))
        ref_p.grad = (grad * grad_scale).double()
        ref32.grad = grad * grad_scale
        opt.step()
        opt32.step()
        ops.call("eavit_adam_step", p, grad.cuda(), m, v, sh, n, step, lr, b1, b2, eps, grad_scale)
        assert int(step.item()) == k + 1
        st = opt.state[ref_p]
        pd, md, vd = p.cpu().double(), m.cpu().double(), v.cpu().double()
        # weights: 1e-6 absolute (|p| < 8: a float32 ulp is up to 4.8e-7); exp_avg mixes signs, so its error is measured
        # against the largest moment, not element-wise (0.9 m + 0.1 g cancels to ~0 for some of 10^6 elements);
        # exp_avg_sq is a sum of positive terms: element-wise relative
        assert float((pd - ref_p.detach()).abs().max()) < 1e-6
        assert float((md - st["exp_avg"]).abs().max() / st["exp_avg"].abs().max()) < 1e-6
        assert float(((vd - st["exp_avg_sq"]).abs() / (st["exp_avg_sq"].abs() + 1e-30)).max()) < 1e-5
        # as close to torch's own float32 Adam as float32 Adam is to the exact recurrence
        assert float((p.cpu() - ref32.detach()).abs().max()) < 1e-6
        s32 = opt32.state[ref32]
        assert float(((v.cpu() - s32["exp_avg_sq"]).abs() / (s32["exp_avg_sq"].abs() + 1e-30)).max()) < 1e-6
        assert torch.equal(sh.cpu(), _bf16_rn(p.cpu()))      # shadow = round-to-nearest-even bf16 of the new master weights


def test_adam_tick_apply_ranges_skip_frozen():
    """eavit_adam_tick + eavit_adam_apply over two ranges == eavit_adam_step on those ranges; the gap keeps p, m, v."""
    from eavit_b200 import ops
    n = 4096
    g = torch.Generator().manual_seed(5)
    p0, grad = torch.randn(n, generator=g), torch.randn(n, generator=g)
    m0, v0 = torch.randn(n, generator=g) * 0.1, torch.rand(n, generator=g) * 0.01
    A = dict(p=p0.clone().cuda(), m=m0.clone().cuda(), v=v0.clone().cuda(), s=torch.zeros(n, dtype=torch.bfloat16, device="cuda"),
             step=torch.full((1,), 4, dtype=torch.int64, device="cuda"))
    Bf = dict(p=p0.clone().cuda(), m=m0.clone().cuda(), v=v0.clone().cuda(), s=torch.zeros(n, dtype=torch.bfloat16, device="cuda"),
              step=torch.full((1,), 4, dtype=torch.int64, device="cuda"))
    gd = grad.cuda()
    ops.call("eavit_adam_step", A["p"], gd, A["m"], A["v"], A["s"], n, A["step"], 1e-3, 0.9, 0.999, 1e-8, 0.5)
    ops.call("eavit_adam_tick", Bf["step"])
    for lo, hi in ((0, 1024), (2048, 4096)):
        ops.call("eavit_adam_apply", Bf["p"][lo:hi], gd[lo:hi], Bf["m"][lo:hi], Bf["v"][lo:hi], Bf["s"][lo:hi], hi - lo, Bf["step"],
                 1e-3, 0.9, 0.999, 1e-8, 0.5)
    assert int(Bf["step"].item()) == 5 == int(A["step"].item())
    for lo, hi in ((0, 1024), (2048, 4096)):
        for k in ("p", "m", "v"):
            assert torch.equal(A[k][lo:hi], Bf[k][lo:hi]), k
    assert torch.equal(Bf["p"][1024:2048].cpu(), p0[1024:2048]) and torch.equal(Bf["m"][1024:2048].cpu(), m0[1024:2048])
    assert torch.equal(Bf["v"][1024:2048].cpu(), v0[1024:2048])


def test_sumsq_and_global_grad_norm_match_reference_formula():
    """utils.global_grad_norm_ (fused sum of squares) vs the reference's per-tensor loop (utils.py:141-170)."""
    from eavit_b200 import ops, utils
    g = torch.Generator().manual_seed(9)
    params = []
    for shp in ((256, 144), (256,), (768, 256), (3, 5, 7), (1024, 256), (18,)):
        p = torch.nn.Parameter(torch.zeros(shp, device="cuda"))
        p.grad = (torch.randn(shp, generator=g) * 0.3).cuda()
        params.append(p)
    params.append(torch.nn.Parameter(torch.zeros(4, device="cuda")))           # no gradient: filtered out (utils.py:158)
    total = 0.0
    for p in params:
        if p.grad is not None:
            total += p.grad.data.double().norm(2).item() ** 2                   # the reference's loop, in float64
    want = total ** 0.5
    got = utils.global_grad_norm_(params)
    assert abs(got - want) <= 1e-5 * want
    assert utils.global_grad_norm_(params[0]) == pytest.approx(params[0].grad.double().norm().item(), rel=1e-5)
    # raw entry point on one flat buffer
    x = torch.randn(1_000_000, generator=g).cuda()
    acc = torch.zeros(1, device="cuda")
    ops.call("eavit_sumsq_f32", x, x.numel(), acc)
    assert float(acc.item()) == pytest.approx(float(x.double().pow(2).sum().item()), rel=1e-5)


@pytest.mark.parametrize("world", [1, 4])
@pytest.mark.parametrize("scale", [0.01, 30.0])
def test_clip_by_norm_matches_clip_grad_norm(world, scale):
    """eavit_sumsq_f32 + eavit_clip_by_norm vs nn.utils.clip_grad_norm_ (agents.py:496-499).  With `world` ranks the flat
    buffer holds the SUM of the ranks' gradients and the threshold is max_norm * world (agents.py here :318-321); after
    Adam's 1 / world this equals clipping the MEAN gradient at max_norm."""
    from eavit_b200 import ops
    n, max_norm = 65536, 0.5
    g = torch.Generator().manual_seed(11)
    mean_grad = torch.randn(n, generator=g) * scale / (n ** 0.5)               # ||g|| ~ scale: both the clipped and the untouched case
    p = torch.nn.Parameter(torch.zeros(n))
    p.grad = mean_grad.clone()
    before = float(torch.nn.utils.clip_grad_norm_([p], max_norm))
    assert (before > max_norm) == (scale > 1)
    summed = (mean_grad * world).cuda()
    nrm = torch.zeros(1, device="cuda")
    ops.call("eavit_sumsq_f32", summed, n, nrm)
    ops.call("eavit_clip_by_norm", summed, n, nrm, max_norm * world)
    got = summed.cpu() / world
    assert float((got - p.grad).norm() / p.grad.norm()) < 1e-6
    assert float(nrm.sqrt().item()) / world == pytest.approx(before, rel=1e-5)


def _small_agent(E=2, T=8, **conf):
    import eavit_b200  # noqa
    from eavit_b200 import agents, config, utils
    cfg = O.OracleConfig(lr=1e-3, epoch=1, mini_batch=2)
    config.load_config(None, ViTlucidrains_dropout=0.0, ViTlucidrains_emb_dropout=0.0, **conf)
    agent = agents.RNDAgent(84, cfg.n_actions, utils.Env_action_space_type.DISCRETE, E, T, cfg.gamma, GAE_Lambda=cfg.lam,
                            learning_rate=cfg.lr, ent_coef=cfg.ent_coef, epoch=cfg.epoch, batch_size=E * T // cfg.mini_batch,
                            ppo_eps=cfg.ppo_eps, use_cuda=True, representation_lr_method="None", device="cuda", logger=utils.Logger())
    P = O.init_params(cfg, seed=7)
    agent.load_state_dict({k: v.clone() for k, v in P.items()}, strict=True)
    roll = O.synth_rollout(E=E, T=T, seed=21)
    args = O.prepare_update(cfg, T, E, roll, O.RunningMeanStd(shape=(1, 1, 84, 84)), O.RunningMeanStd(), O.RewardForwardFilter(cfg.int_gamma))
    return agent, P, cfg, args


def test_train_step_update_equals_torch_adam_on_the_same_gradient():
    """The optimiser part of an update in isolation: take the flat gradient the kernels produced (apply=False), feed the
    SAME gradient to torch.optim.Adam on a copy of the weights and to the agent's fused optimiser: identical to 1e-6 per
    step over 3 steps (moments carried)."""
    agent, P, cfg, args = _small_agent()
    R = agent.upload_rollout(*args)
    rt = agent.runtime()
    st = rt.store
    names = list(st.shapes)
    ref = [torch.nn.Parameter(st.w(n).detach().cpu().clone()) for n in names]
    opt = torch.optim.Adam(ref, lr=cfg.lr)
    B = 8
    for k in range(3):
        idx = torch.arange(B, device="cuda") + (k % 2) * B
        mask = torch.tensor((np.arange(B) % 2).astype(np.float32)).cuda()
        agent.train_step(R, idx, mask, None, apply=False)
        for n, p in zip(names, ref):
            p.grad = st.g(n).detach().cpu().clone()
        opt.step()
        # the fused update on the very same gradient buffer (a second train_step would re-accumulate the split-K weight
        # gradients with red.add in another order: bitwise different noise-level gradients, which Adam normalises to +-lr)
        agent.optimizer.step()
        for n, p in zip(names, ref):
            d = float((st.w(n).detach().cpu() - p.detach()).abs().max())
            assert d < 1e-6, (k, n, d)
            assert torch.equal(st.b16(n).cpu(), st.w(n).cpu().to(torch.bfloat16)), n
            p.data.copy_(st.w(n).detach().cpu())                  # keep both trajectories on the same weights


def test_frozen_backbone_is_not_trained():
    """train.py:261-263: requires_grad = False on model.feature.* -> torch Adam never touches those tensors (weights,
    moments), clip / grad-norm ignore them; heads and the RND predictor still train."""
    agent, P, cfg, args = _small_agent()
    R = agent.upload_rollout(*args)
    B = 8
    idx, mask = torch.arange(B, device="cuda"), torch.ones(B, device="cuda")
    agent.train_step(R, idx, mask, None, apply=True)              # one ordinary step: non-zero moments everywhere
    rt = agent.runtime()
    st = rt.store
    for p in agent.model.feature.parameters():
        p.requires_grad = False
    snap = {n: (st.w(n).clone(), st._view(st.m, n).clone(), st._view(st.v, n).clone()) for n in st.shapes}
    agent.train_step(R, idx + B, mask, None, apply=True)
    moved = 0
    for n in st.shapes:
        w0, m0, v0 = snap[n]
        if n.startswith("model.feature."):
            assert torch.equal(st.w(n), w0) and torch.equal(st._view(st.m, n), m0) and torch.equal(st._view(st.v, n), v0), n
            assert float(st.g(n).abs().max()) == 0.0, n
        elif float((st.w(n) - w0).abs().max()) > 0:
            moved += 1
    assert moved >= 10
    assert int(st.step.item()) == 2
    # unfreezing resumes training of the backbone
    for p in agent.model.feature.parameters():
        p.requires_grad = True
    w_before = st.w("model.feature.transformer.layers.0.0.to_qkv.weight").clone()
    agent.train_step(R, idx, mask, None, apply=True)
    assert float((st.w("model.feature.transformer.layers.0.0.to_qkv.weight") - w_before).abs().max()) > 0


def test_optimizer_state_dict_is_torch_adam_layout():
    """agent.optimizer.state_dict() (checkpoint entry of train.py:931) loads into a real torch.optim.Adam over the same
    tensors and continues identically; a torch.optim.Adam state_dict loads back into the fused optimiser."""
    agent, P, cfg, args = _small_agent()
    R = agent.upload_rollout(*args)
    B = 8
    idx, mask = torch.arange(B, device="cuda"), torch.ones(B, device="cuda")
    for _ in range(2):
        agent.train_step(R, idx, mask, None, apply=True)
    st = agent.runtime().store
    names = list(st.shapes)
    sd = agent.optimizer.state_dict()
    assert set(sd) >= {"state", "param_groups"} and sd["param_groups"][0]["params"] == list(range(len(names)))
    assert sd["eavit_param_names"] == names
    # -> torch
    ref = [torch.nn.Parameter(st.w(n).detach().clone()) for n in names]
    opt = torch.optim.Adam(ref, lr=123.0)
    opt.load_state_dict({"state": sd["state"], "param_groups": sd["param_groups"]})
    assert opt.param_groups[0]["lr"] == pytest.approx(cfg.lr)
    agent.train_step(R, idx, mask, None, apply=False)
    for n, p in zip(names, ref):
        p.grad = st.g(n).detach().clone()
    opt.step()
    agent.optimizer.step()                                        # the fused update on the same gradient buffer
    for n, p in zip(names, ref):
        assert float((st.w(n) - p.detach()).abs().max()) < 1e-6, n
    # -> back (from torch's own state_dict, which carries no names: same index order here, unique by construction)
    tsd = opt.state_dict()
    tsd["eavit_param_names"] = names
    m_before = st.m.clone()
    st.m.zero_(); st.v.zero_(); st.step.zero_()
    agent.optimizer.load_state_dict(tsd)
    assert int(st.step.item()) == 3
    assert float((st.m - m_before).abs().max()) < 1e-6
    # names that do not belong to this agent are refused instead of scrambling the moments
    bad = dict(tsd)
    bad["eavit_param_names"] = ["nope." + n for n in names]
    with pytest.raises(ValueError):
        agent.optimizer.load_state_dict(bad)


def test_stale_autograd_backward_is_refused():
    """The autograd path keeps activations in per-batch scratch buffers: a second forward of the same batch size before
    backward would silently corrupt the gradients -- it must raise instead."""
    agent, P, cfg, args = _small_agent()
    x = torch.rand(4, 4, 84, 84, device="cuda")
    pol, ve, vi = agent.model(x)
    with torch.no_grad():
        agent.model(x * 0.5)                                       # overwrites the batch-4 activations
    with pytest.raises(RuntimeError, match="overwritten"):
        (pol.sum() + ve.sum()).backward()
    pol, ve, vi = agent.model(x)
    with torch.no_grad():
        agent.model(torch.rand(6, 4, 84, 84, device="cuda"))      # another batch size: separate buffers, fine
    (pol.sum() + ve.sum()).backward()
