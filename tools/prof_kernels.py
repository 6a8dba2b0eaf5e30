"""Launch a few instances of the hot kernels at the cfg3 shapes (for `ncu -k regex:...`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import eavit_b200
from eavit_b200 import ops

torch.manual_seed(0)
T, D, MLP, H, DH = 201216, 256, 1024, 8, 32
which = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = 3
if which in ("all", "gemm"):
    x = torch.randn(T, D, device="cuda").bfloat16()
    w1 = (torch.randn(MLP, D, device="cuda") / 16).bfloat16()
    b1 = torch.randn(MLP, device="cuda")
    hact = torch.empty(T, MLP, device="cuda", dtype=torch.bfloat16)
    hpre = torch.empty_like(hact)
    for _ in range(reps):
        ops.gemm(x, w1, bias=b1, act=ops.ACT_GELU, out_bf16=hact, out_pre=hpre)          # MLP1 fwd
    w2 = (torch.randn(D, MLP, device="cuda") / 32).bfloat16()
    dh = torch.empty_like(hact)
    for _ in range(reps):
        ops.gemm(x, w2, b_mn=True, act=ops.ACT_GELU_BWD, aux=hpre, out_bf16=dh)            # MLP2 dX + GELU'
    res = torch.randn(T, D, device="cuda")
    out = torch.empty_like(res)
    for _ in range(reps):
        ops.gemm(hact, w2, bias=torch.randn(D, device="cuda"), residual=res, out_f32=out)  # MLP2 fwd + residual
    wq = (torch.randn(3 * D, D, device="cuda") / 16).bfloat16()
    qkv = torch.empty(T, 3 * D, device="cuda", dtype=torch.bfloat16)
    for _ in range(reps):
        ops.gemm(x, wq, out_bf16=qkv)                                                       # QKV
    dW = torch.zeros(MLP, D, device="cuda")
    for _ in range(reps):
        ops.gemm(dh, x, a_mn=True, b_mn=True, out_f32=dW, atomic=True, split_k=10)          # dW1 split-K
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
    for _ in range(reps):
        ops.call("eavit_attention_fwd_tc", qkv, ss, len(lens), 197, qkv.shape[0], H, DH, DH ** -0.5, o, lse, 0.0, 0)
    for _ in range(reps):
        ops.call("eavit_attention_bwd_tc", qkv, do, lse, ss, len(lens), 197, qkv.shape[0], H, DH, DH ** -0.5, dqkv, 0.0, 0)
        ops.call("eavit_attention_bwd_tct", qkv, o, do, lse, ss, len(lens), 197, qkv.shape[0], H, DH, DH ** -0.5, dqkv, 0.0, 0)
torch.cuda.synchronize()
print("done")
