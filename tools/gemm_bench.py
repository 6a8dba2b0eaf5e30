"""CUDA-event timing of the block GEMMs at the cfg3 shapes (T = 201 216 token rows), one line per epilogue variant."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import eavit_b200
from eavit_b200 import ops

torch.manual_seed(0)
T, D, MLP = 201216, 256, 1024
x = torch.randn(T, D, device="cuda").bfloat16()
w1 = (torch.randn(MLP, D, device="cuda") / 16).bfloat16()
b1 = torch.randn(MLP, device="cuda")
hact = torch.empty(T, MLP, device="cuda", dtype=torch.bfloat16)
hpre = torch.empty_like(hact)
w2 = (torch.randn(D, MLP, device="cuda") / 32).bfloat16()
dh = torch.empty_like(hact)
cs = torch.zeros(MLP, device="cuda")
res = torch.randn(T, D, device="cuda")
out = torch.empty_like(res)
xn = torch.empty(T, D, device="cuda", dtype=torch.bfloat16)
g, be = torch.ones(D, device="cuda"), torch.zeros(D, device="cuda")
mu, rs = torch.empty(T, device="cuda"), torch.empty(T, device="cuda")
b2 = torch.randn(D, device="cuda")
wo = (torch.randn(D, D, device="cuda") / 16).bfloat16()
wq = (torch.randn(3 * D, D, device="cuda") / 16).bfloat16()
qkv = torch.empty(T, 3 * D, device="cuda", dtype=torch.bfloat16)
from eavit_b200.engine import _split_k
dWq = torch.zeros(768, 256, device="cuda"); dW1 = torch.zeros(1024, 256, device="cuda"); dW2 = torch.zeros(256, 1024, device="cuda"); dWo = torch.zeros(256, 256, device="cuda")
cases = {
    "mlp1_fwd_gelu_save_grad  <256,6>": lambda: ops.gemm(x, w1, bias=b1, act=ops.ACT_GELU_SAVE_GRAD, out_bf16=hact, out_pre=hpre),
    "mlp2_dx_mul_aux+colsum   <256,7>": lambda: ops.gemm(x, w2, b_mn=True, act=ops.ACT_MUL_AUX, aux=hpre, out_bf16=dh, colsum=cs),
    "mlp2_fwd_resid           <256,4>": lambda: ops.gemm(hact, w2, bias=b2, residual=res, out_f32=out),
    "mlp2_fwd_resid_ln        <256,8>": lambda: ops.gemm(hact, w2, bias=b2, residual=res, out_f32=out, out_bf16=xn, ln=(g, be, mu, rs, 1e-5)),
    "outproj_fwd_resid_ln K=256 <256,8>": lambda: ops.gemm(x, wo, bias=b2, residual=res, out_f32=out, out_bf16=xn, ln=(g, be, mu, rs, 1e-5)),
    "qkv_fwd_store            <256,1>": lambda: ops.gemm(x, wq, out_bf16=qkv),
    "dW_qkv  M=768 N=256 splitk": lambda: ops.gemm(qkv, x, a_mn=True, b_mn=True, out_f32=dWq, atomic=True, split_k=_split_k(768, 256, T)),
    "dW_mlp1 M=1024 N=256 splitk": lambda: ops.gemm(hact, x, a_mn=True, b_mn=True, out_f32=dW1, atomic=True, split_k=_split_k(1024, 256, T)),
    "dW_mlp2 M=256 N=1024 splitk": lambda: ops.gemm(x, hact, a_mn=True, b_mn=True, out_f32=dW2, atomic=True, split_k=_split_k(256, 1024, T)),
    "dW_out  M=256 N=256 splitk": lambda: ops.gemm(x, xn, a_mn=True, b_mn=True, out_f32=dWo, atomic=True, split_k=_split_k(256, 256, T)),
    "qkv_dx  N=256 K=768 kmn bf16 out": lambda: ops.gemm(qkv, wq, b_mn=True, out_bf16=xn),
    "mlp1_dx N=256 K=1024 kmn bf16 out": lambda: ops.gemm(hact, w1, b_mn=True, out_bf16=xn),
    "outproj_dx N=256 K=256 kmn bf16 out": lambda: ops.gemm(x, wo, b_mn=True, out_bf16=xn),
}
H, DH, B = 8, 32, 512
lens = [196] * B + [197] * B
st = [0]
for n_ in lens:
    st.append(st[-1] + n_)
ss = torch.tensor(st, dtype=torch.int32, device="cuda")
aqkv = torch.randn(st[-1], 3 * H * DH, device="cuda").bfloat16()
ao = torch.empty(st[-1], H * DH, device="cuda", dtype=torch.bfloat16)
lse = torch.empty(st[-1], H, device="cuda")
ado = torch.randn_like(ao)
adqkv = torch.empty_like(aqkv)
cases["attention_fwd_tc"] = lambda: ops.call("eavit_attention_fwd_tc", aqkv, ss, len(lens), 197, aqkv.shape[0], H, DH, DH ** -0.5, ao, lse, 0.0, 0)
cases["attention_bwd_tct"] = lambda: ops.call("eavit_attention_bwd_tct", aqkv, ao, ado, lse, ss, len(lens), 197, aqkv.shape[0], H, DH, DH ** -0.5, adqkv, 0.0, 0)
cases["attention_bwd_tc (query-major)"] = lambda: ops.call("eavit_attention_bwd_tc", aqkv, ado, lse, ss, len(lens), 197, aqkv.shape[0], H, DH, DH ** -0.5, adqkv, 0.0, 0)
sel = sys.argv[1:] or None
for name, fn in cases.items():
    if sel and not any(s in name for s in sel):
        continue
    for _ in range(3):
        fn()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
    ev[0].record()
    for i in range(10):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(10))
    print(f"{name:38s} median {ts[5]:7.1f} us   min {ts[0]:7.1f} us")
if sel and "ksweep" in sel:
    # same output bytes, shrinking K: separates the operand side (TMA / L2 -> SM traffic, MMA) from the epilogue + HBM writes
    for K in (256, 128, 64):
        xa = torch.randn(T, K, device="cuda").bfloat16()
        wk = (torch.randn(3 * D, K, device="cuda") / 16).bfloat16()
        w1k = (torch.randn(MLP, K, device="cuda") / 16).bfloat16()
        for name, fn in ((f"qkv store K={K}", lambda: ops.gemm(xa, wk, out_bf16=qkv)),
                         (f"mlp1 gelu save K={K}", lambda: ops.gemm(xa, w1k, bias=b1, act=ops.ACT_GELU_SAVE_GRAD, out_bf16=hact, out_pre=hpre))):
            for _ in range(3):
                fn()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
            ev[0].record()
            for i in range(10):
                fn()
                ev[i + 1].record()
            torch.cuda.synchronize()
            ts = sorted(ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(10))
            print(f"{name:38s} median {ts[5]:7.1f} us   min {ts[0]:7.1f} us")
