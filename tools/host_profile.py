"""Host-side cost of one optimiser step (python + ctypes + tensor-map encodes): cProfile over 32 eager steps."""
import cProfile, pstats, os, sys, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
torch.cuda.set_device(0)
agent = bench.make_agent(bench.E_PER_GPU)
upd_args, u8 = bench.synth_update_args(bench.E_PER_GPU, 0)
R = agent.upload_rollout(*upd_args)
B = agent.batch_size
perm = torch.randperm(R["states"].shape[0], device="cuda")
mask = (torch.rand(B, device="cuda") < 0.25).float()
for i in range(4):
    agent.train_step(R, perm[B * i: B * (i + 1)], mask)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for i in range(32):
    agent.train_step(R, perm[B * (i % 30): B * (i % 30 + 1)], mask)
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(18)
print(s.getvalue()[:4000])
