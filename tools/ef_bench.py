import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import os, torch, numpy as np
from test_gpu_model import CFGS, make_agent
from oracle import oracle as O
agent,P=make_agent(O.OracleConfig(),2,8)
rt=agent.runtime()
B=512
for dt in (torch.uint8, torch.float32):
    img=torch.randint(0,256,(2048,4,84,84),dtype=torch.uint8,device='cuda')
    if dt==torch.float32: img=img.float()/255
    idx=torch.randperm(2048,device='cuda')[:B]
    enc=rt.encoder
    # time only the embedding: call forward pieces via env toggle and profile table
    from eavit_b200 import ops
    for fused in ("1","0"):
        os.environ["EAVIT_FUSE_EMBED"]=fused
        for _ in range(2): enc.forward(img,B,idx)
        ops.profile_start()
        for _ in range(3): enc.forward(img,B,idx)
        tab=ops.profile_stop()
        keys=[k for k in tab if any(s in k for s in ("embed","patchify","K=144")) or k=="layernorm_fwd"]
        tot=sum(tab[k][1]/3 for k in keys if k!="layernorm_fwd")
        print(dt, "fused" if fused=="1" else "4-launch", {k: round(tab[k][1]/3*1e3,1) for k in keys}, "sum (w/o LN) us", round(tot*1e3,1))
