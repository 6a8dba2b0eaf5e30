"""GPU parity: RND convolution (de)materialisation kernels vs torch unfold / conv-transpose arithmetic (model.py:368-416)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from eavit_b200 import ops as _ops
    return _ops


CONVS = [(84, 1, 8, 4), (20, 32, 4, 2), (9, 64, 3, 1)]      # (H, C, k, stride) of the three RND convolutions


@pytest.mark.parametrize("H,C,k,s", CONVS)
@pytest.mark.parametrize("split3", [0, 1])
def test_im2col_matches_unfold(ops, H, C, k, s, split3):
    torch.manual_seed(H + C)
    N, B = 7, 5
    x = torch.randn(N, H, H, C, device="cuda")                       # NHWC fp32
    idx = torch.tensor([6, 0, 3, 3, 1], dtype=torch.int64, device="cuda")
    OH = (H - k) // s + 1
    K = C * k * k
    col = torch.full((B * OH * OH, (3 if split3 else 1) * K), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.call("eavit_im2col", x, ops.F32, idx, B, H, H, C, k, k, s, col, split3)
    torch.cuda.synchronize()
    ref = F.unfold(x[idx].permute(0, 3, 1, 2), k, stride=s)         # [B, C*k*k, L], K order (c, ki, kj) like Conv2d weights
    ref = ref.permute(0, 2, 1).reshape(B * OH * OH, K)
    hi = ref.bfloat16()
    assert torch.equal(col[:, :K], hi)
    if split3:
        assert torch.equal(col[:, K:2 * K], hi)
        lo = (ref - hi.float()).bfloat16()
        assert torch.equal(col[:, 2 * K:], lo)


@pytest.mark.parametrize("H,C,k,s", CONVS[1:])
def test_col2im_lrelu_matches_fold(ops, H, C, k, s):
    torch.manual_seed(H * 3 + C)
    B = 4
    OH = (H - k) // s + 1
    K = C * k * k
    dcol = torch.randn(B * OH * OH, K, device="cuda").bfloat16()
    act = torch.randn(B, H, H, C, device="cuda").bfloat16()
    din = torch.full((B * H * H, C), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.call("eavit_col2im_lrelu", dcol, act, B, H, H, C, k, k, s, din)
    torch.cuda.synchronize()
    folded = F.fold(dcol.float().reshape(B, OH * OH, K).permute(0, 2, 1), (H, H), k, stride=s)   # [B, C, H, H]
    ref = folded.permute(0, 2, 3, 1) * torch.where(act.float() > 0, 1.0, 0.01)
    err = ((din.float().reshape(B, H, H, C) - ref).norm() / ref.norm()).item()
    assert torch.isfinite(din.float()).all()
    assert err < 4e-3, err            # bf16 rounding of the output only (fp32 tap sums)
