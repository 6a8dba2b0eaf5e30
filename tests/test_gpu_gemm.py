"""GPU parity: the tcgen05 GEMM (C ABI eavit_gemm_bf16) vs torch fp32 matmul on the same bf16 operands."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from eavit_b200 import ops as _ops
    assert torch.cuda.is_available()
    return _ops


def rel_err(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-20)).item()


SHAPES = [(128, 256, 64), (256, 768, 256), (1000, 256, 1024), (393, 1024, 256), (130, 64, 144), (64, 32, 64),
          (4096, 768, 256), (77, 512, 3136)]


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_tn_plain(ops, M, N, K):
    torch.manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda").bfloat16()
    B = torch.randn(N, K, device="cuda").bfloat16()
    ref = A.float() @ B.float().t()
    out = torch.empty(M, N, device="cuda", dtype=torch.float32)
    out16 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(A, B, out_f32=out, out_bf16=out16)
    torch.cuda.synchronize()
    assert rel_err(out, ref) < 1e-5, rel_err(out, ref)          # fp32 accumulate of identical bf16 operands
    assert rel_err(out16, ref) < 5e-3


@pytest.mark.parametrize("a_mn,b_mn", [(False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", [(256, 256, 512), (768, 256, 1000), (200, 1024, 256), (256, 144, 776)])
def test_gemm_mn_major_operands(ops, a_mn, b_mn, M, N, K):
    """dX = dY W (B stored [K,N]) and dW = dY^T X (both operands stored [K, *])."""
    torch.manual_seed(1)
    A = torch.randn(M, K, device="cuda").bfloat16()
    B = torch.randn(N, K, device="cuda").bfloat16()
    ref = A.float() @ B.float().t()
    Ain = A.t().contiguous() if a_mn else A
    Bin = B.t().contiguous() if b_mn else B
    out = torch.empty(M, N, device="cuda", dtype=torch.float32)
    ops.gemm(Ain, Bin, a_mn=a_mn, b_mn=b_mn, out_f32=out)
    torch.cuda.synchronize()
    assert rel_err(out, ref) < 1e-5, rel_err(out, ref)


def test_gemm_split_k_atomic(ops):
    torch.manual_seed(2)
    M, N, K = 768, 256, 5000
    A = torch.randn(K, M, device="cuda").bfloat16()
    B = torch.randn(K, N, device="cuda").bfloat16()
    ref = A.float().t() @ B.float()
    out = torch.zeros(M, N, device="cuda", dtype=torch.float32)
    ops.gemm(A, B, a_mn=True, b_mn=True, out_f32=out, atomic=True, split_k=16)
    torch.cuda.synchronize()
    assert rel_err(out, ref) < 1e-5
    ops.gemm(A, B, a_mn=True, b_mn=True, out_f32=out, atomic=True, split_k=7)   # accumulates on top
    torch.cuda.synchronize()
    assert rel_err(out, 2 * ref) < 1e-5


@pytest.mark.parametrize("M,N,K,mn,split", [(768, 256, 5000, True, 16), (1024, 256, 9000, True, 37), (256, 1024, 4100, True, 9),
                                             (704, 512, 3000, True, 5), (384, 256, 2048, False, 4), (333, 264, 1000, False, 1)])
def test_gemm_split_k_paired_tiles(ops, M, N, K, mn, split):
    """Weight-gradient GEMMs take their 128-row tiles in pairs (both TMEM accumulators per work item): odd tile counts, M / N
    tails, MN-major and K-major operands, accumulation on top of earlier content."""
    torch.manual_seed(M + K)
    At = torch.randn(K, M, device="cuda").bfloat16()
    Bt = torch.randn(K, N, device="cuda").bfloat16()
    ref = At.float().t() @ Bt.float()
    out = torch.full((M + 2, N), 3.0, device="cuda", dtype=torch.float32)
    if mn:
        ops.gemm(At, Bt, a_mn=True, b_mn=True, out_f32=out[:M], atomic=True, split_k=split)
    else:
        ops.gemm(At.t().contiguous(), Bt.t().contiguous(), out_f32=out[:M], atomic=True, split_k=split)
    torch.cuda.synchronize()
    assert rel_err(out[:M], ref + 3.0) < 2e-5, rel_err(out[:M], ref + 3.0)
    assert bool((out[M:] == 3.0).all())


def test_gemm_epilogues(ops):
    torch.manual_seed(3)
    M, N, K = 500, 1024, 256
    A = torch.randn(M, K, device="cuda").bfloat16()
    B = (torch.randn(N, K, device="cuda") / 16).bfloat16()
    bias = torch.randn(N, device="cuda")
    res = torch.randn(M, N, device="cuda")
    pre_ref = A.float() @ B.float().t() + bias
    # bias + GELU, saving the pre-activation (MLP1 forward, vit.py:29-30)
    out16 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    pre16 = torch.empty_like(out16)
    ops.gemm(A, B, bias=bias, act=ops.ACT_GELU, out_bf16=out16, out_pre=pre16)
    assert rel_err(pre16, pre_ref) < 5e-3
    assert rel_err(out16, torch.nn.functional.gelu(pre_ref)) < 5e-3
    # bias + residual -> fp32 (attention out-proj / MLP2, vit.py:88-89)
    out = torch.empty(M, N, device="cuda", dtype=torch.float32)
    ops.gemm(A, B, bias=bias, residual=res, out_f32=out)
    assert rel_err(out, pre_ref + res) < 1e-5
    # in-place residual
    res2 = res.clone()
    ops.gemm(A, B, bias=bias, residual=res2, out_f32=res2)
    assert rel_err(res2, pre_ref + res) < 1e-5
    # GELU backward epilogue: v * gelu'(aux)
    aux = torch.randn(M, N, device="cuda").bfloat16()
    x = aux.float().requires_grad_(True)
    torch.nn.functional.gelu(x).sum().backward()
    ops.gemm(A, B, act=ops.ACT_GELU_BWD, aux=aux, out_bf16=out16)
    assert rel_err(out16, (A.float() @ B.float().t()) * x.grad) < 5e-3
    # the pair the engine uses: the forward stores gelu'(pre) next to gelu(pre), the backward multiplies (vit.py:27-37)
    g16 = torch.empty_like(out16)
    ops.gemm(A, B, bias=bias, act=ops.ACT_GELU_SAVE_GRAD, out_bf16=out16, out_pre=g16)
    xr = pre_ref.clone().requires_grad_(True)
    torch.nn.functional.gelu(xr).sum().backward()
    assert rel_err(out16, torch.nn.functional.gelu(pre_ref)) < 5e-3
    assert rel_err(g16, xr.grad) < 5e-3
    dh16 = torch.empty_like(out16)
    cs = torch.zeros(N, device="cuda")
    ops.gemm(A, B, act=ops.ACT_MUL_AUX, aux=g16, out_bf16=dh16, colsum=cs)
    ref_dh = (A.float() @ B.float().t()) * g16.float()
    assert rel_err(dh16, ref_dh) < 5e-3
    assert rel_err(cs, ref_dh.sum(0)) < 2e-3
    # LeakyReLU / ReLU and their backward masks
    ops.gemm(A, B, bias=bias, act=ops.ACT_LRELU, out_f32=out)
    assert rel_err(out, torch.nn.functional.leaky_relu(pre_ref)) < 1e-5
    ops.gemm(A, B, bias=bias, act=ops.ACT_RELU, out_f32=out)
    assert rel_err(out, torch.relu(pre_ref)) < 1e-5
    ops.gemm(A, B, act=ops.ACT_RELU_BWD, aux=aux, out_f32=out)
    assert rel_err(out, (A.float() @ B.float().t()) * (aux.float() > 0)) < 1e-5
    ops.gemm(A, B, act=ops.ACT_LRELU_BWD, aux=aux, out_f32=out)
    assert rel_err(out, (A.float() @ B.float().t()) * torch.where(aux.float() > 0, 1.0, 0.01)) < 1e-5
    torch.cuda.synchronize()


@pytest.mark.parametrize("M,N,K", [(500, 1024, 256), (333, 144, 256), (4100, 256, 1024), (40000, 1024, 256), (700, 4096, 64)])
def test_gemm_fused_column_sums(ops, M, N, K):
    """colsum[n] += sum_m of the stored values (bias gradient fused into the dX epilogue); accumulates across calls."""
    torch.manual_seed(4)
    A = torch.randn(M, K, device="cuda").bfloat16()
    B = (torch.randn(N, K, device="cuda") / 16).bfloat16()
    aux = torch.randn(M, N, device="cuda").bfloat16()
    x = aux.float().requires_grad_(True)
    torch.nn.functional.gelu(x).sum().backward()
    ref = (A.float() @ B.float().t()) * x.grad
    out16 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    cs = torch.zeros(N, device="cuda")
    ops.gemm(A, B, act=ops.ACT_GELU_BWD, aux=aux, out_bf16=out16, colsum=cs)
    assert rel_err(out16, ref) < 5e-3
    assert rel_err(cs, ref.sum(0)) < 2e-3, rel_err(cs, ref.sum(0))
    out = torch.empty(M, N, device="cuda")
    ops.gemm(A, B, out_f32=out, colsum=cs)                    # generic epilogue, accumulates on top
    assert rel_err(cs, ref.sum(0) + (A.float() @ B.float().t()).sum(0)) < 2e-3
    torch.cuda.synchronize()


@pytest.mark.parametrize("M,N,K,b_mn", [(500, 1024, 256, False), (333, 320, 128, False), (40000, 1024, 256, True), (20000, 768, 256, False),
                                        (19000, 448, 64, True), (4100, 256, 1024, True), (129, 264, 512, False)])
def test_gemm_tma_epilogues(ops, M, N, K, b_mn):
    """The epilogues that work in the accumulator's native layout and move boxes through the TMA unit (plain bf16 store with
    the weight-stationary operand ring for K <= 256; multiply-by-aux + column sums): M / N tails clipped by the tensor map,
    column blocks in which a warp owns no box, many tiles per CTA, MN-major B, repeated launches accumulating the sums."""
    torch.manual_seed(M + N)
    A = torch.randn(M, K, device="cuda").bfloat16()
    B = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
    Bop = B.t().contiguous() if b_mn else B
    ref = A.float() @ B.float().t()
    out16 = torch.full((M + 3, N), 7.0, device="cuda", dtype=torch.bfloat16)        # guard rows behind the M tail
    ops.gemm(A, Bop, b_mn=b_mn, out_bf16=out16[:M])
    assert rel_err(out16[:M], ref) < 5e-3
    assert bool((out16[M:] == 7.0).all())
    bias0 = torch.randn(N, device="cuda")
    ops.gemm(A, Bop, b_mn=b_mn, bias=bias0, out_bf16=out16[:M])                           # + bias (HF-style qkv)
    assert rel_err(out16[:M], ref + bias0) < 5e-3
    aux = torch.randn(M, N, device="cuda").bfloat16()
    cs = torch.zeros(N, device="cuda")
    dh = torch.full((M + 3, N), 7.0, device="cuda", dtype=torch.bfloat16)
    for _ in range(2):
        ops.gemm(A, Bop, b_mn=b_mn, act=ops.ACT_MUL_AUX, aux=aux, out_bf16=dh[:M], colsum=cs)
    want = ref * aux.float()
    assert rel_err(dh[:M], want) < 5e-3
    assert bool((dh[M:] == 7.0).all())
    assert rel_err(cs, 2 * want.sum(0)) < 3e-3, rel_err(cs, 2 * want.sum(0))
    ops.gemm(A, Bop, b_mn=b_mn, act=ops.ACT_MUL_AUX, aux=aux, out_bf16=dh[:M])           # without the column sums
    assert rel_err(dh[:M], want) < 5e-3
    # bias + GELU, storing gelu'(pre) next to gelu(pre)
    bias = torch.randn(N, device="cuda")
    h = torch.full((M + 3, N), 7.0, device="cuda", dtype=torch.bfloat16)
    g = torch.full((M + 3, N), 7.0, device="cuda", dtype=torch.bfloat16)
    ops.gemm(A, Bop, b_mn=b_mn, bias=bias, act=ops.ACT_GELU_SAVE_GRAD, out_bf16=h[:M], out_pre=g[:M])
    pre = (ref + bias).requires_grad_(True)
    hr = torch.nn.functional.gelu(pre)
    hr.sum().backward()
    assert rel_err(h[:M], hr.detach()) < 5e-3
    assert rel_err(g[:M], pre.grad) < 5e-3
    assert bool((h[M:] == 7.0).all()) and bool((g[M:] == 7.0).all())
    torch.cuda.synchronize()


def test_gemm_rejects_bad_arguments(ops):
    A = torch.randn(64, 60, device="cuda").bfloat16()     # pitch 120 B, not 16-byte aligned
    B = torch.randn(64, 60, device="cuda").bfloat16()
    out = torch.empty(64, 64, device="cuda")
    with pytest.raises(RuntimeError):
        ops.gemm(A, B, out_f32=out)


@pytest.mark.parametrize("M,K", [(500, 256), (4100, 1024), (128, 64), (1, 256)])
def test_gemm_residual_epilogue_with_fused_layernorm(ops, M, K):
    """x' = residual + A W^T + b (fp32) and LayerNorm(x') * gamma + beta (bf16) + (mean, rstd) from ONE launch (N = 256 = one
    tile per row): vit.py:88-89 residual adds followed by the PreNorm LayerNorm of vit.py:28 / :47."""
    torch.manual_seed(M + K)
    N = 256
    A = torch.randn(M, K, device="cuda").bfloat16()
    B = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
    bias = torch.randn(N, device="cuda")
    res = torch.randn(M, N, device="cuda") * 3 + 0.7
    gamma, beta = torch.randn(N, device="cuda") * 0.2 + 1, torch.randn(N, device="cuda") * 0.1
    out = torch.empty(M, N, device="cuda")
    xn = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    mean, rstd = torch.empty(M, device="cuda"), torch.empty(M, device="cuda")
    ops.gemm(A, B, bias=bias, residual=res, out_f32=out, out_bf16=xn, ln=(gamma, beta, mean, rstd, 1e-5))
    ref = A.float() @ B.float().t() + bias + res
    assert rel_err(out, ref) < 1e-5
    ln_ref = torch.nn.functional.layer_norm(out, (N,), gamma, beta, 1e-5)          # of the values the kernel itself stored
    assert rel_err(xn, ln_ref) < 4e-3                                               # bf16 rounding of the output
    assert torch.allclose(mean, out.mean(1), rtol=1e-4, atol=1e-5)
    assert torch.allclose(rstd, (out.var(1, unbiased=False) + 1e-5).rsqrt(), rtol=1e-4)
    # the statistics need not be written
    xn2 = torch.empty_like(xn)
    ops.gemm(A, B, bias=bias, residual=res, out_f32=out, out_bf16=xn2, ln=(gamma, beta, None, None, 1e-5))
    assert torch.equal(xn2, xn)
    with pytest.raises(RuntimeError):                                               # needs N == 256
        ops.gemm(A, B[:128], bias=bias[:128], residual=res[:, :128].contiguous(), out_f32=out[:, :128].contiguous(),
                 out_bf16=xn[:, :128].contiguous(), ln=(gamma[:128], beta[:128], None, None, 1e-5))
