"""Micro-benchmarks of the hot kernels at the cfg3 shapes (CUDA events, L2 flushed between launches by rotating buffers).

    python tools/bench_kernels.py [attn|gemm|ln|all]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import eavit_b200  # noqa: F401
from eavit_b200 import ops

torch.manual_seed(0)
T, D, MLP, H, DH = 201216, 256, 1024, 8, 32
which = sys.argv[1] if len(sys.argv) > 1 else "all"
REPS = 10


def timed(name, fn, flops=None, bytes_=None):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(REPS):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / REPS
    extra = ""
    if flops:
        extra += f"  {flops / ms / 1e9:8.1f} TFLOP/s"
    if bytes_:
        extra += f"  {bytes_ / ms / 1e6:8.1f} GB/s"
    print(f"{name:44s} {ms * 1e3:9.1f} us{extra}", flush=True)


if which in ("all", "attn"):
    B = 512
    lens = [196] * B + [197] * B
    st = [0]
    for n in lens:
        st.append(st[-1] + n)
    ss = torch.tensor(st, dtype=torch.int32, device="cuda")
    qkv = torch.randn(st[-1], 3 * H * DH, device="cuda").bfloat16()
    o = torch.empty(st[-1], H * DH, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(st[-1], H, device="cuda")
    do = torch.randn_like(o)
    dqkv = torch.empty_like(qkv)
    fl = sum(4.0 * n * n * DH * H for n in lens)
    timed("attention_fwd_tc  (1024 seq x 8 heads)", lambda: ops.call("eavit_attention_fwd_tc", qkv, ss, len(lens), 197, qkv.shape[0], H, DH, DH ** -0.5, o, lse, 0.0, 0), fl,
          qkv.numel() * 2 + o.numel() * 2)
    timed("attention_bwd_tc", lambda: ops.call("eavit_attention_bwd_tc", qkv, do, lse, ss, len(lens), 197, qkv.shape[0], H, DH, DH ** -0.5, dqkv, 0.0, 0), 2.5 * fl,
          2 * qkv.numel() * 2 + o.numel() * 2)
    timed("attention_bwd_tct (+ D pre-kernel)", lambda: ops.call("eavit_attention_bwd_tct", qkv, o, do, lse, ss, len(lens), 197, qkv.shape[0], H, DH, DH ** -0.5, dqkv, 0.0, 0), 2.5 * fl,
          2 * qkv.numel() * 2 + 3 * o.numel() * 2)

if which in ("all", "gemm"):
    x = torch.randn(T, D, device="cuda").bfloat16()
    w1 = (torch.randn(MLP, D, device="cuda") / 16).bfloat16()
    b1 = torch.randn(MLP, device="cuda")
    hact = torch.empty(T, MLP, device="cuda", dtype=torch.bfloat16)
    hpre = torch.empty_like(hact)
    timed("MLP1 fwd + GELU   [T,256]x[1024,256]", lambda: ops.gemm(x, w1, bias=b1, act=ops.ACT_GELU, out_bf16=hact, out_pre=hpre), 2.0 * T * D * MLP,
          T * D * 2 + 2 * T * MLP * 2)
    w2 = (torch.randn(D, MLP, device="cuda") / 32).bfloat16()
    dh = torch.empty_like(hact)
    cs = torch.zeros(MLP, device="cuda")
    timed("MLP2 dX + GELU'   [T,256]x[256,1024]", lambda: ops.gemm(x, w2, b_mn=True, act=ops.ACT_GELU_BWD, aux=hpre, out_bf16=dh, colsum=cs),
          2.0 * T * D * MLP, T * D * 2 + 2 * T * MLP * 2)
    res = torch.randn(T, D, device="cuda")
    out = torch.empty_like(res)
    b2 = torch.randn(D, device="cuda")
    timed("MLP2 fwd + resid  [T,1024]x[256,1024]", lambda: ops.gemm(hact, w2, bias=b2, residual=res, out_f32=out), 2.0 * T * D * MLP,
          T * MLP * 2 + 2 * T * D * 4)
    wo = (torch.randn(D, D, device="cuda") / 16).bfloat16()
    timed("out-proj + resid  [T,256]x[256,256]", lambda: ops.gemm(x, wo, bias=b2, residual=res, out_f32=out), 2.0 * T * D * D, T * D * 2 + 2 * T * D * 4)
    wq = (torch.randn(3 * D, D, device="cuda") / 16).bfloat16()
    qkv2 = torch.empty(T, 3 * D, device="cuda", dtype=torch.bfloat16)
    timed("QKV               [T,256]x[768,256]", lambda: ops.gemm(x, wq, out_bf16=qkv2), 2.0 * T * D * 3 * D, T * D * 2 + T * 3 * D * 2)
    dxn = torch.empty(T, D, device="cuda", dtype=torch.bfloat16)
    timed("MLP1 dX           [T,1024]x[1024,256]", lambda: ops.gemm(dh, w1, b_mn=True, out_bf16=dxn), 2.0 * T * D * MLP, T * MLP * 2 + T * D * 2)
    dW = torch.zeros(MLP, D, device="cuda")
    timed("dW1 split-K       [1024,T]x[T,256]", lambda: ops.gemm(dh, x, a_mn=True, b_mn=True, out_f32=dW, atomic=True, split_k=18), 2.0 * T * D * MLP,
          T * MLP * 2 + T * D * 2)

if which in ("all", "ln"):
    F32, BF16 = ops.F32, ops.BF16
    x = torch.randn(T, D, device="cuda")
    g = torch.randn(D, device="cuda")
    b = torch.randn(D, device="cuda")
    y = torch.empty(T, D, device="cuda", dtype=torch.bfloat16)
    m = torch.empty(T, device="cuda")
    r = torch.empty(T, device="cuda")
    timed("layernorm_fwd f32 -> bf16", lambda: ops.call("eavit_layernorm_fwd", x, D, g, b, y, BF16, D, m, r, T, D, 1e-5), None, T * D * 6)
    dy = torch.randn(T, D, device="cuda").bfloat16()
    dres = torch.randn(T, D, device="cuda")
    dx = torch.empty(T, D, device="cuda")
    dx16 = torch.empty(T, D, device="cuda", dtype=torch.bfloat16)
    dg, db, dsum = torch.zeros(D, device="cuda"), torch.zeros(D, device="cuda"), torch.zeros(D, device="cuda")
    timed("layernorm_bwd (dy bf16, dres, dx f32+bf16)",
          lambda: ops.call("eavit_layernorm_bwd", dy, BF16, D, x, D, m, r, g, dres, D, dx, D, dx16, D, dg, db, dsum, 0.0, 0, T, D), None, T * D * (2 + 4 + 4 + 4 + 2))
print("done")
