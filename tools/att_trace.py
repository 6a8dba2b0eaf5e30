"""Phase timing of the attention backward kernel (needs the -DEAVIT_TRACE build: EAVIT_B200_LIB=tools/_trace/libeavit_b200_trace.so)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import eavit_b200  # noqa
from eavit_b200 import ops, _lib
H, DH, B = 8, 32, 512
lens = [196] * B + [197] * B
if os.environ.get("ATT_TRACE_SHAPE") == "hg":          # vit_hg at the cfg4 per-GPU shape: 1024 sequences of 50 tokens, 16 heads x 64
    H, DH, B = 16, 64, 512
    lens = [50] * (2 * B)
st = [0]
for n in lens:
    st.append(st[-1] + n)
ss = torch.tensor(st, dtype=torch.int32, device="cuda")
qkv = torch.randn(st[-1], 3 * H * DH, device="cuda").bfloat16()
o = torch.empty(st[-1], H * DH, device="cuda", dtype=torch.bfloat16)
lse = torch.empty(st[-1], H, device="cuda")
do = torch.randn_like(o)
dqkv = torch.empty_like(qkv)
ops.call("eavit_attention_fwd_tc", qkv, ss, len(lens), max(lens), qkv.shape[0], H, DH, DH ** -0.5, o, lse, 0.0, 0)
which = sys.argv[1] if len(sys.argv) > 1 else "bwd"
def run():
    if which == "bwdt":
        ops.call("eavit_attention_bwd_tct", qkv, o, do, lse, ss, len(lens), max(lens), qkv.shape[0], H, DH, DH ** -0.5, dqkv, 0.0, 0)
    elif which == "bwd":
        ops.call("eavit_attention_bwd_tc", qkv, do, lse, ss, len(lens), max(lens), qkv.shape[0], H, DH, DH ** -0.5, dqkv, 0.0, 0)
    else:
        ops.call("eavit_attention_fwd_tc", qkv, ss, len(lens), max(lens), qkv.shape[0], H, DH, DH ** -0.5, o, lse, 0.0, 0)
run()
torch.cuda.synchronize()
L = _lib.lib()
buf = (ctypes.c_longlong * 32)()
dbg = L.eavit_debug_bt_trace if which == "bwdt" else L.eavit_debug_att_trace
dbg(buf, 1)
run()
dbg(buf, 0)
v = list(buf)
tot = sum(v)
names = sys.argv[2].split(",") if len(sys.argv) > 2 else [str(i) for i in range(32)]
for i, x in enumerate(v):
    if x:
        print(f"{i:2d} {names[i] if i < len(names) else '':28s} {x:12d} cycles {100.0 * x / tot:5.1f}%")
print("total cycles (block 0)", tot)
