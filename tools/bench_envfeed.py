"""Per-step cost of the learner-side env exchange with synthetic workers that replay pre-generated
frames (no ALE in this image): the reference's receive loop + float32 conversion (train.py:605, :615-654) vs
StepCollector over the same pipe protocol vs StepCollector with FrameRing workers.

    python tools/bench_envfeed.py [E] [steps] [cuda]
"""
import multiprocessing as mp
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch


def worker(conn, seed, steps, ring_name, env_idx, num_env):
    rng = np.random.default_rng(seed)
    ring = None
    if ring_name is not None:
        from eavit_b200.envfeed import FrameRing
        ring = FrameRing.attach(ring_name, num_env)
    frames = rng.integers(0, 256, (8, 4, 84, 84)).astype(np.float64)      # replayed: generation is not what is measured
    conn.send(frames[0])
    for t in range(steps):
        conn.recv()
        s = frames[t % 8]
        if ring is not None:
            ring.write(t, env_idx, s)
            conn.send([None, 0.0, False, False, {}])
        else:
            conn.send([s, 0.0, False, False, {}])
    conn.close()


def spawn(E, steps, ring_name=None):
    ctx = mp.get_context("spawn")
    conns, procs = [], []
    for i in range(E):
        a, b = ctx.Pipe()
        p = ctx.Process(target=worker, args=(b, i, steps, ring_name, i, E), daemon=True)
        p.start()
        conns.append(a)
        procs.append(p)
    return conns, procs


def main():
    E = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    dev = "cuda" if len(sys.argv) > 3 and sys.argv[3] == "cuda" else None
    import eavit_b200  # noqa
    from eavit_b200.envfeed import FrameRing, StepCollector
    out = {}
    # (a) reference loop
    conns, procs = spawn(E, steps)
    [c.recv() for c in conns]
    t0 = None
    for t in range(steps):
        if t == 5:
            t0 = time.perf_counter()
        for c in conns:
            c.send(0)
        ns = np.zeros([E, 4, 84, 84], dtype=np.float64)
        no = np.zeros([E, 1, 84, 84], dtype=np.float64)
        for i, c in enumerate(conns):
            s, r, d, tr, vr = c.recv()
            ns[i] = s[:]
            no[i] = s[3].reshape(1, 84, 84)
        x = torch.from_numpy(np.float32(ns) / 255.)                   # what get_action receives (train.py:605)
        if dev:
            x = x.to(dev); torch.cuda.synchronize()
    out["reference loop (float64 pipe messages + float32 conversion)"] = (time.perf_counter() - t0) / (steps - 5)
    [p.join(timeout=10) for p in procs]
    # (b) collector, same protocol
    conns, procs = spawn(E, steps)
    col = StepCollector(conns, device=dev)
    col.initial_states()
    for t in range(steps):
        if t == 5:
            t0 = time.perf_counter()
        got = col.step([0] * E)
        if dev:
            torch.cuda.synchronize()
    out["StepCollector, reference pipe protocol (uint8 staging)"] = (time.perf_counter() - t0) / (steps - 5)
    [p.join(timeout=10) for p in procs]
    # (c) collector + shared-memory ring workers
    ring = FrameRing(E)
    conns, procs = spawn(E, steps, ring.name)
    col = StepCollector(conns, device=dev, ring=ring)
    col.initial_states()
    for t in range(steps):
        if t == 5:
            t0 = time.perf_counter()
        got = col.step([0] * E)
        if dev:
            torch.cuda.synchronize()
    out["StepCollector + FrameRing workers (frames in shared memory)"] = (time.perf_counter() - t0) / (steps - 5)
    [p.join(timeout=10) for p in procs]
    ring.close()
    print(f"E = {E} synthetic env workers, {os.cpu_count()} host cores, device = {dev}")
    for k, v in out.items():
        print(f"  {k:66s} {v * 1e3:8.2f} ms/step  {E / v:9.0f} env-steps/s")


if __name__ == "__main__":
    main()
