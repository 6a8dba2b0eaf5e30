"""GPU parity: attention kernels (CUDA-core v1 and tcgen05) vs torch fp32 softmax attention on the same bf16 inputs."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from eavit_b200 import ops as _ops
    return _ops


def ref_attention(qkv, starts, H, Dh, scale):
    T = qkv.shape[0]
    out = torch.zeros(T, H * Dh, device=qkv.device)
    lse = torch.zeros(T, H, device=qkv.device)
    q, k, v = qkv.float().split(H * Dh, dim=1)
    for s0, s1 in zip(starts[:-1], starts[1:]):
        for h in range(H):
            sl = slice(h * Dh, (h + 1) * Dh)
            sc = (q[s0:s1, sl] @ k[s0:s1, sl].t()) * scale
            out[s0:s1, sl] = sc.softmax(-1) @ v[s0:s1, sl]
            lse[s0:s1, h] = torch.logsumexp(sc, -1)
    return out, lse


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


CASES = [([196, 197, 196, 197], 8, 32), ([50, 50, 50], 2, 64), ([197] * 5, 8, 32), ([1, 17, 128, 129, 224], 4, 32),
         ([64, 200], 3, 64), ([197, 33, 196, 5] * 40, 8, 32),        # > 148 items, every CTA loops
         ([50] * 9, 4, 64),                    # equal even lengths: packed 4 (fwd) / 2 (bwd) per work item, ragged last pack
         ([64] * 7, 4, 32),                    # packed 3 per item, two query tiles
         ([16] * 30, 2, 32),                   # packed 14 per item
         ([129, 144, 150, 176, 192, 208], 2, 32)]   # two key tiles with every k-step count 9..13 (transposed-score backward)


@pytest.mark.parametrize("lens,H,Dh", CASES)
@pytest.mark.parametrize("impl", ["v1", "tc"])
def test_attention_forward(ops, lens, H, Dh, impl):
    torch.manual_seed(sum(lens) + H)
    starts = [0]
    for n in lens:
        starts.append(starts[-1] + n)
    T = starts[-1]
    qkv = (torch.randn(T, 3 * H * Dh, device="cuda") * 1.5).bfloat16()
    ss = torch.tensor(starts, dtype=torch.int32, device="cuda")
    out = torch.full((T, H * Dh), float("nan"), device="cuda", dtype=torch.bfloat16)
    lse = torch.full((T, H), float("nan"), device="cuda")
    scale = Dh ** -0.5
    if impl == "v1":
        ops.call("eavit_attention_fwd", qkv, ss, len(lens), max(lens), H, Dh, scale, out, lse)
    else:
        ops.call("eavit_attention_fwd_tc", qkv, ss, len(lens), max(lens), T, H, Dh, scale, out, lse, 0.0, 0)
    torch.cuda.synchronize()
    ro, rl = ref_attention(qkv, starts, H, Dh, scale)
    assert torch.isfinite(out.float()).all() and torch.isfinite(lse).all()
    assert rel(out, ro) < 6e-3, rel(out, ro)
    assert (lse - rl).abs().max().item() < 2e-2


BWD_CASES = [([196, 197, 196, 197], 8, 32), ([50, 50, 50], 2, 64), ([1, 17, 128, 129, 224], 4, 32), ([64, 100], 3, 64),
             ([197, 33, 196, 5] * 40, 8, 32), ([50] * 9, 4, 64), ([64] * 7, 4, 32), ([16] * 30, 2, 32),
             ([129, 144, 150, 176, 192, 208], 2, 32), ([1, 17, 128, 129, 208, 2, 31, 160], 4, 32)]


@pytest.mark.parametrize("lens,H,Dh", BWD_CASES)
@pytest.mark.parametrize("impl", ["v1", "tc", "tct"])
def test_attention_backward(ops, lens, H, Dh, impl):
    if impl == "tct" and not ((Dh == 32 and max(lens) <= 208 and H % 2 == 0) or (Dh == 64 and max(lens) <= 128)):
        pytest.skip("transposed-score backward: Dh = 32 up to 208 tokens, Dh = 64 up to 128")
    torch.manual_seed(sum(lens) + H + 1)
    starts = [0]
    for n in lens:
        starts.append(starts[-1] + n)
    T = starts[-1]
    qkv = (torch.randn(T, 3 * H * Dh, device="cuda") * 1.2).bfloat16()
    dout = torch.randn(T, H * Dh, device="cuda").bfloat16()
    ss = torch.tensor(starts, dtype=torch.int32, device="cuda")
    scale = Dh ** -0.5
    x = qkv.float().requires_grad_(True)
    ro, _ = ref_attention_autograd(x, starts, H, Dh, scale)
    (ro * dout.float()).sum().backward()
    out = torch.empty(T, H * Dh, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(T, H, device="cuda")
    dqkv = torch.full((T, 3 * H * Dh), float("nan"), device="cuda", dtype=torch.bfloat16)
    if impl == "tct":
        ops.call("eavit_attention_fwd_tc", qkv, ss, len(lens), max(lens), T, H, Dh, scale, out, lse, 0.0, 0)
        ops.call("eavit_attention_bwd_tct", qkv, out, dout, lse, ss, len(lens), max(lens), T, H, Dh, scale, dqkv, 0.0, 0)
    elif impl == "tc":
        ops.call("eavit_attention_fwd_tc", qkv, ss, len(lens), max(lens), T, H, Dh, scale, out, lse, 0.0, 0)
        ops.call("eavit_attention_bwd_tc", qkv, dout, lse, ss, len(lens), max(lens), T, H, Dh, scale, dqkv, 0.0, 0)
    else:
        ops.call("eavit_attention_fwd", qkv, ss, len(lens), max(lens), H, Dh, scale, out, lse)
        ops.call("eavit_attention_bwd", qkv, out, dout, lse, ss, len(lens), max(lens), H, Dh, scale, dqkv)
    torch.cuda.synchronize()
    assert torch.isfinite(dqkv.float()).all()
    g = x.grad
    n = H * Dh
    for name, sl in (("dq", slice(0, n)), ("dk", slice(n, 2 * n)), ("dv", slice(2 * n, 3 * n))):
        e = rel(dqkv[:, sl], g[:, sl])
        assert e < 1.2e-2, (name, e)


def ref_attention_autograd(x, starts, H, Dh, scale):
    q, k, v = x.split(H * Dh, dim=1)
    outs = []
    for s0, s1 in zip(starts[:-1], starts[1:]):
        hs = []
        for h in range(H):
            sl = slice(h * Dh, (h + 1) * Dh)
            sc = (q[s0:s1, sl] @ k[s0:s1, sl].t()) * scale
            hs.append(sc.softmax(-1) @ v[s0:s1, sl])
        outs.append(torch.cat(hs, dim=1))
    return torch.cat(outs, dim=0), None


@pytest.mark.parametrize("lens,H,Dh", [([196, 197, 196, 197], 8, 32), ([50, 50, 50], 2, 64), ([1, 17, 128, 129, 224, 300], 4, 32)])
@pytest.mark.parametrize("p", [0.0, 0.1])
def test_attention_pooled_row_only(ops, lens, H, Dh, p):
    """Last-layer attention for query row 0 of each sequence: forward equals row 0 of full attention, backward equals the
    full backward fed with a gradient that is zero outside row 0 (dq zero elsewhere, dk / dv dense)."""
    torch.manual_seed(sum(lens) + 7)
    seed = ops.site_seed(3, 5)
    starts = [0]
    for n in lens:
        starts.append(starts[-1] + n)
    T, nseq = starts[-1], len(lens)
    qkv = (torch.randn(T, 3 * H * Dh, device="cuda") * 1.2).bfloat16()
    ss = torch.tensor(starts, dtype=torch.int32, device="cuda")
    scale = Dh ** -0.5
    x = qkv.float().requires_grad_(True)
    q, k, v = x.split(H * Dh, dim=1)
    rows = []
    for s0, s1 in zip(starts[:-1], starts[1:]):
        hs = []
        for h in range(H):
            sl = slice(h * Dh, (h + 1) * Dh)
            pr = ((q[s0:s0 + 1, sl] @ k[s0:s1, sl].t()) * scale).softmax(-1)
            if p > 0:
                pr = pr * ops.dropout_mask(1, s1 - s0, p, seed, row0=s0, col0=h * 256)
            hs.append(pr @ v[s0:s1, sl])
        rows.append(torch.cat(hs, dim=1))
    ref = torch.cat(rows, dim=0)                                  # [nseq, H*Dh]
    d0 = torch.randn(nseq, H * Dh, device="cuda").bfloat16()
    (ref * d0.float()).sum().backward()
    out0 = torch.full((nseq, H * Dh), float("nan"), device="cuda", dtype=torch.bfloat16)
    dqkv = torch.full((T, 3 * H * Dh), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.call("eavit_attention_row0_fwd", qkv, ss, nseq, max(lens), H, Dh, scale, out0, p, seed)
    ops.call("eavit_attention_row0_bwd", qkv, d0, ss, nseq, max(lens), H, Dh, scale, dqkv, p, seed)
    torch.cuda.synchronize()
    assert rel(out0, ref) < 6e-3, rel(out0, ref)
    assert torch.isfinite(dqkv.float()).all()
    n = H * Dh
    for name, sl in (("dq", slice(0, n)), ("dk", slice(n, 2 * n)), ("dv", slice(2 * n, 3 * n))):
        e = rel(dqkv[:, sl], x.grad[:, sl])
        assert e < 8e-3, (name, e)
    not_first = torch.ones(T, dtype=torch.bool, device="cuda")
    not_first[torch.tensor(starts[:-1], device="cuda")] = False
    assert (dqkv[not_first][:, :n] == 0).all()                    # exact zeros for query rows >= 1
