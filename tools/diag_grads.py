"""GPU diagnostic: per-tensor gradient / forward errors of the kernels vs the CPU oracle (prints a table)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import oracle as O
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_model import make_agent, CFGS, _batch, rel

which = sys.argv[1] if len(sys.argv) > 1 else "lucid"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
cfg = CFGS[which]
E, T = 2, 16
agent, P = make_agent(cfg, E, T)
args = _batch(cfg, E, T)
states, te, ti, y, adv, obs, old = args
N = E * T
idx = np.random.default_rng(1).permutation(N)[:B]
mask = (np.random.default_rng(2).random(B) < 0.5).astype(np.float32)
for k in O.trainable_names(P):
    P[k].requires_grad_(True)
old_flat = torch.tensor(old).permute(1, 0, 2).contiguous().view(-1, cfg.n_actions)
ti_ = torch.from_numpy(idx)
loss, terms, (pol_o, ve_o, vi_o) = O.ppo_rnd_loss(
    P, cfg, torch.FloatTensor(states)[ti_], torch.FloatTensor(te)[ti_], torch.FloatTensor(ti)[ti_], torch.LongTensor(y)[ti_],
    torch.FloatTensor(adv)[ti_], torch.FloatTensor(obs)[ti_], old_flat[ti_], torch.tensor(mask))
loss.backward()
R = agent.upload_rollout(*args)
stats = torch.zeros(16, device="cuda")
agent.train_step(R, torch.from_numpy(idx).cuda(), torch.tensor(mask).cuda(), stats, apply=False)
s = stats.cpu().numpy()
print("terms oracle", terms)
print("terms kernel", dict(actor=s[1], critic_ext=s[2], critic_int=s[3], entropy=s[4], rnd=s[5]))
rt = agent.runtime()
w = agent._ws[(B, cfg.n_actions)]
hb = rt.heads.buf[B]
print("fwd policy rel", rel(hb.t["policy"].cpu().numpy(), pol_o.detach().numpy()),
      "v rel", rel(hb.t["v"].cpu().numpy(), torch.cat((vi_o, ve_o)).detach().reshape(-1).numpy()))
st = rt.store
groups = {"model": ([], []), "rnd": ([], [])}
rows = []
for k in O.trainable_names(P):
    g_ref = P[k].grad
    g = st.g(k).cpu()
    if g_ref is None:
        rows.append((k, "none", float(g.abs().max()))); continue
    gr = groups["model" if k.startswith("model.") else "rnd"]
    if not k.endswith("attention.key.bias"):
        gr[0].append(g.reshape(-1).numpy()); gr[1].append(g_ref.reshape(-1).numpy())
    rows.append((k, rel(g.numpy(), g_ref.numpy()), float(g_ref.norm())))
for r in rows:
    print("%-70s %-12s %.3e" % (r[0], ("%.4f" % r[1]) if not isinstance(r[1], str) else r[1], r[2]))
for name, (a, b) in groups.items():
    print(name, "total rel", rel(np.concatenate(a), np.concatenate(b)), "norm", float(np.linalg.norm(np.concatenate(b))))
