"""eavit_sgemm_small: the fp32 CUDA-core GEMM of the policy / value heads (model.py:227-246 forward, their dX / dW)."""
import pytest
import torch

from eavit_b200 import ops

pytestmark = pytest.mark.gpu


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("M,N,K", [(512, 256, 256), (1024, 448, 256), (130, 18, 448), (64, 64, 31), (333, 70, 1000)])
@pytest.mark.parametrize("tA,tB", [(0, 0), (1, 1), (0, 1), (1, 0)])
def test_sgemm_small_all_modes(M, N, K, tA, tB):
    torch.manual_seed(M + N + K + 2 * tA + tB)
    A = torch.randn(M, K, device="cuda")
    B = torch.randn(N, K, device="cuda") / K ** 0.5
    bias = torch.randn(N, device="cuda")
    Ast = A.t().contiguous() if tA else A                      # transA: stored [K, M]
    Bst = B.t().contiguous() if tB else B                      # transB: stored [K, N]
    ref = A.double() @ B.double().t()
    C = torch.empty(M, N, device="cuda")
    # bias + ReLU (forward layer)
    ops.call("eavit_sgemm_small", Ast, Ast.stride(0), tA, Bst, Bst.stride(0), tB, bias, None, C, N, M, N, K, 1, 0)
    assert rel(C, torch.relu(ref + bias.double())) < 2e-6
    # plain, then accumulate on top (weight gradients: split-K with atomics when the tile count is small)
    ops.call("eavit_sgemm_small", Ast, Ast.stride(0), tA, Bst, Bst.stride(0), tB, None, None, C, N, M, N, K, 0, 0)
    assert rel(C, ref) < 2e-6
    ops.call("eavit_sgemm_small", Ast, Ast.stride(0), tA, Bst, Bst.stride(0), tB, None, None, C, N, M, N, K, 0, 1)
    assert rel(C, 2 * ref) < 2e-6
    # ReLU-backward mask: zero where the saved activation is not positive
    aux = torch.randn(M, N, device="cuda")
    ops.call("eavit_sgemm_small", Ast, Ast.stride(0), tA, Bst, Bst.stride(0), tB, None, aux, C, N, M, N, K, 0, 0)
    assert rel(C, ref * (aux > 0)) < 2e-6
    torch.cuda.synchronize()
