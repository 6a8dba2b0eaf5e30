import os, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import bench
torch.cuda.set_device(0)
agent = bench.make_agent(bench.E_PER_GPU)
upd_args, u8 = bench.synth_update_args(bench.E_PER_GPU, 0)
agent.epoch = 4
def t(fn, n=1):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): r = fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n, r
agent.train_model(*upd_args, 0)            # warm
dt_all, _ = t(lambda: agent.train_model(*upd_args, 1))
dt_up, R = t(lambda: agent.upload_rollout(*upd_args))
dev_args = (R["states"], R["te"].double(), R["ti"].double(), R["y"], R["adv"].double(), R["obs"], R["old"])
dt_dev, _ = t(lambda: agent.train_model(*dev_args, 2))
print(f"train_model(host args) {dt_all*1e3:.1f} ms; upload_rollout alone {dt_up*1e3:.1f} ms; train_model(device args) {dt_dev*1e3:.1f} ms")
# host-side only: how long does the python loop take to ENQUEUE 128 steps (no sync)?
torch.cuda.synchronize(); t0 = time.perf_counter(); agent.train_model(*dev_args, 3); t_enq = time.perf_counter() - t0; torch.cuda.synchronize()
print(f"host enqueue time of one update (device args, no sync): {t_enq*1e3:.1f} ms")
# true host cost per step: enqueue 4 steps into an EMPTY launch queue (no back-pressure from the GPU)
B = agent.batch_size
perm = torch.randperm(R["states"].shape[0], device="cuda")
mask = (torch.rand(B, device="cuda") < 0.25).float()
best = 1e9
for rep in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(4):
        agent.train_step(R, perm[B * i: B * (i + 1)], mask)
    best = min(best, (time.perf_counter() - t0) / 4)
    torch.cuda.synchronize()
print(f"host cost of one step with an empty launch queue: {best*1e3:.2f} ms (GPU time per step ~7.9 ms)")
