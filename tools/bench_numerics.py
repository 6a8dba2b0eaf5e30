"""Achieved HBM bandwidth of the rollout-side numerics kernels at the cfg3 / cfg5 sizes.

    python tools/bench_numerics.py [E]     (E envs x 128 steps; default 128 and 1024)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import eavit_b200  # noqa: F401
from eavit_b200 import ops

T, F = 128, 84 * 84


def timed(name, fn, nbytes, reps=10, flush=None):
    fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        if flush is not None:
            flush.zero_()                      # > L2: the next launch reads from HBM
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    ms = tot / reps
    print(f"{name:58s} {ms * 1e3:9.1f} us  {nbytes / 1e6:9.1f} MB  {nbytes / ms / 1e6:8.1f} GB/s", flush=True)


for E in ([int(a) for a in sys.argv[1:]] or [128, 1024]):
    N = E * T
    print(f"--- E = {E} envs x {T} steps (N = {N} observations of 84x84)")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    mean = torch.zeros(F, dtype=torch.float64, device="cuda")
    var = torch.ones(F, dtype=torch.float64, device="cuda")
    cnt = torch.full((1,), 1e-4, dtype=torch.float64, device="cuda")
    for dt, sz in ((torch.uint8, 1), (torch.float32, 4)):
        x = (torch.rand(N, F, device="cuda") * 255).to(dt)
        timed(f"rms_update      x {str(dt)[6:]:8s} -> mean/var f64", lambda: ops.rms_update(x, mean, var, cnt), N * F * sz, flush=flush)
        out = torch.empty(N, F, dtype=torch.float32, device="cuda")
        timed(f"obs_normalize   x {str(dt)[6:]:8s} -> f32", lambda: ops.obs_normalize(x, mean, var, out=out), N * F * (sz + 4), flush=flush)
        o16 = torch.empty(N, F, dtype=torch.bfloat16, device="cuda")
        timed(f"obs_normalize   x {str(dt)[6:]:8s} -> bf16", lambda: ops.obs_normalize(x, mean, var, out=o16), N * F * (sz + 2), flush=flush)
        del x, out, o16
    r = torch.rand(E, T, device="cuda")
    v = torch.randn(E, T + 1, device="cuda")
    timed("gae_f32 (warp-shuffle scan, one warp per env)", lambda: ops.gae_f32(r, None, v, 0.99, 0.95), E * T * 16)
    r64 = r.double()
    done = (torch.rand(E, T, device="cuda") < 0.05).to(torch.uint8)
    timed("gae_f64 (numpy-promotion exact)", lambda: ops.gae_f64(r64, done, v, 0.999, 0.95, 0), E * T * (8 + 1 + 4 + 16))
    rew = torch.zeros(E, device="cuda")
    timed("reward_filter + moments", lambda: ops.reward_filter(r, rew, True, 0.99), E * T * 4)
    a, b = torch.randn(N, 512, device="cuda"), torch.randn(N, 512, device="cuda")
    timed("intrinsic_mse [N,512] x2 -> [N]", lambda: ops.intrinsic_mse(a, b), N * 512 * 8, flush=flush)
print("done")
