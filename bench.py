#!/usr/bin/env python
"""Headline benchmark: PPO+RND update samples/sec of the ViT explorative-attention RND agent on B200.

    python bench.py --gpus N --steps K --warmup W          # this repo (sm_100a kernels through the C ABI)
    python bench.py --impl reference --steps K --warmup W   # the reference's CPU path (oracle port), host cores

Workload (BASELINE.json configs[2], SURVEY 8d cfg3 -- the configuration the metric is quoted on): lucidrains
explorative-attention ViT (dim 256, depth 3, 8x32 heads, mlp 1024, patch 6 -> 196/197 tokens) + PPO heads + RND
predictor/target, 128 envs x 128 steps per GPU (N = 16 384 samples), MiniBatch 32 -> 512 samples per optimiser step,
dropout keys = 0.0 (the parity configuration).

A "step" = one minibatch optimiser step (agents.py:284-508): batch gather, RND fwd/bwd, ViT fwd/bwd (both
attention passes), heads, PPO/RND loss, [gradient all-reduce], Adam.   value = steps * 512 * n_gpus / time.
`value` is timed with the rollout resident in HBM; `e2e` times the reference-facing `RNDAgent.train_model(...)`
call with HOST numpy buffers (pinned) -- H2D of the whole rollout + all Epoch x MiniBatch steps + D2H of the stats.
Multi-GPU: weak scaling, every rank owns 128 envs, NCCL all-reduce of the flat gradient per step.

Beside the headline the JSON line carries the other BASELINE configs as records (each measured live by this command):
`cfg2_cnn_backbone` (configs[1]: original RND CNN backbone, 64 envs x 128 steps), `vit_hg_cfg4` (configs[3]: HF-style ViT
1024/12L/16h, 256 envs over the N GPUs of the run), `numerics` (RunningMeanStd update / observation normalisation / GAE
achieved GB/s against the measured copy peak), `sustained` (the same step over >= 200 back-to-back steps) and
`e2e_device_rollout` (DeviceRollout.finish() -> train_model with CUDA tensors).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

E_PER_GPU, T, A = 128, 128, 18
MINI_BATCH, EPOCH = 32, 4
FLOPS_PER_SAMPLE = 6.36e9 + 58.2e6      # ViT+heads fwd+bwd + RND training, per sample (NOMINAL:
                                        # every token of every layer, as the reference computes it)


def executed_fraction(D=256, inner=256, mlp=1024, depth=3, s_a=196, s_b=197):
    """Share of the nominal ViT FLOPs the engine executes: the last layer's out-projection, MLP and attention rows are
    computed for the pooled token only (the reference computes and discards the other 391 of 393 tokens per sample)."""
    qkv, proj, ff = 2 * D * 3 * inner, 2 * inner * D, 4 * D * mlp
    tot = dead = 0.0
    for S in (s_a, s_b):
        per_tok = qkv + proj + ff + 4 * S * inner
        tot += depth * S * per_tok
        dead += (S - 1) * (proj + ff + 4 * S * inner)
    return 1.0 - dead / tot
# dram__bytes_read.sum + dram__bytes_write.sum per launch at the cfg3 shapes (ncu --set full, profiles/r2_ncu_full_top_kernels.md)
NCU_DRAM_BYTES_PER_LAUNCH = {
    "attention_bwd_tct": 794.0e6, "attention_bwd_tc": 683.0e6, "attention_fwd_tc": 395.9e6,
    "gemm_bf16_tcgen05 M=201216 N=1024 K=256 kmn act=7": 892.0e6, "gemm_bf16_tcgen05 M=201216 N=1024 K=256 kk act=8": 875.0e6,
    "gemm_bf16_tcgen05 M=201216 N=256 K=1024 kk act=0 +ln": 895.0e6,
}
WORKLOAD = ("cfg3: lucidrains explorative-attention ViT (expGlados3) RND agent, 128 envs x 128 steps per GPU, "
            "minibatch 512, 4 epochs, dropout keys 0.0")


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.idx, self.samples, self.stop_flag = gpu_index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.idx)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(s[1]) for s in self.samples)
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][2]), "reasons": sorted(reasons),
                "power_w_max": max(float(s[3]) for s in self.samples), "samples": len(sm)}


def synth_update_args(E, seed):
    """Synthetic rollout in the reference's train_model argument dtypes/layouts, in pinned host memory."""
    rng = np.random.default_rng(seed)
    N = E * T

    def pinned(shape, dtype):
        t = torch.empty(shape, dtype=dtype)
        try:
            t = t.pin_memory()
        except Exception:
            pass
        return t
    states = pinned((N, 4, 84, 84), torch.float32)
    u8 = rng.integers(0, 256, (N, 4, 84, 84), dtype=np.uint8)
    np.divide(u8, np.float32(255.0), out=states.numpy(), dtype=np.float32)        # np.float32(total_state) / 255.
    obs = pinned((N, 1, 84, 84), torch.float64)
    obs.numpy()[...] = rng.normal(0, 1, (N, 1, 84, 84)).clip(-5, 5)
    te = rng.normal(0, 1, N)
    ti = rng.normal(0, 1, N)
    adv = rng.normal(0, 1, N)
    y = rng.integers(0, A, N).astype(np.int64)
    old = rng.normal(0, 1, (T, E, A)).astype(np.float32)
    return (states.numpy(), te, ti, y, adv, obs.numpy(), old), u8


def make_agent(E):
    import eavit_b200  # noqa: F401
    from eavit_b200 import agents, config, utils
    conf = os.path.join(ROOT, "configs", "expGlados3_lucidrains_explorative.conf")
    config.load_config(conf, ViTlucidrains_dropout=0.0, ViTlucidrains_emb_dropout=0.0)
    c = config.default_config
    N = E * T
    utils.set_seed(42)
    agent = agents.RNDAgent(84, A, utils.Env_action_space_type.DISCRETE, E, T, float(c["Gamma"]), GAE_Lambda=float(c["GAELambda"]),
                            learning_rate=float(c["LearningRate"]), ent_coef=float(c["Entropy"]), epoch=int(c["Epoch"]),
                            batch_size=int(N / int(c["MiniBatch"])), ppo_eps=float(c["PPOEps"]), use_cuda=True,
                            representation_lr_method="None", device=f"cuda:{torch.cuda.current_device()}", logger=utils.Logger())
    return agent


def cpu_reference(steps, warmup, batch=512, threads=None, rollout=4096):
    """The reference's CPU path for the same optimiser step, on the box's host cores: the oracle port of agents.py:284-508
    (torch fp32, all host threads) at the BENCHED minibatch (512 samples), including what the reference does per minibatch
    before the model runs -- it re-materialises the whole rollout as torch tensors and then indexes the minibatch
    (agents.py:288-301: ``torch.FloatTensor(states)[batch_indices]`` for every array).  The bounded sample is the rollout
    length that re-materialisation runs over (`rollout` samples instead of the 16 384 of cfg3, i.e. a quarter of the
    reference's own per-minibatch copy cost); the model work per step is exactly the benched one.  The 512-sample backward
    is evaluated in 4 chunks of 128 (identical loss and gradient, see oracle.ppo_rnd_backward_chunked) to bound autograd
    memory."""
    from oracle import oracle as O
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    cfg = O.OracleConfig()
    P = O.init_params(cfg, seed=7)
    names = O.trainable_names(P)
    for k in names:
        P[k].requires_grad_(True)
    opt = torch.optim.Adam([P[k] for k in names], lr=cfg.lr)
    rng = np.random.default_rng(0)
    n = batch
    N = max(rollout, n)
    states = np.float32(rng.integers(0, 256, (N, 4, 84, 84), dtype=np.uint8)) / np.float32(255.0)       # train.py:854 dtype
    obs = rng.normal(0, 1, (N, 1, 84, 84)).clip(-5, 5)                                                  # float64, train.py:855
    te, ti, adv = rng.normal(0, 1, N), rng.normal(0, 1, N), rng.normal(0, 1, N)                         # float64
    y = rng.integers(0, A, N).astype(np.int64)
    old = rng.normal(0, 1, (N // 16, 16, A)).astype(np.float32)                                         # [T, E, A]
    sample_range = np.arange(N)
    t_remat = [0.0]

    def step():
        np.random.shuffle(sample_range)
        bi = sample_range[:n]
        t0 = time.perf_counter()
        s_b = torch.FloatTensor(states)[bi]                                                             # agents.py:288
        te_b, ti_b = torch.FloatTensor(te)[bi], torch.FloatTensor(ti)[bi]                               # :289-291
        y_b, adv_b = torch.LongTensor(y)[bi], torch.FloatTensor(adv)[bi]                                # :293-296
        obs_b = torch.FloatTensor(obs)[bi]                                                              # :298
        old_b = torch.tensor(old).permute(1, 0, 2).contiguous().view(-1, A)[bi]                         # :301
        t_remat[0] += time.perf_counter() - t0
        mask = (torch.rand(n) < cfg.update_proportion).float()
        opt.zero_grad()
        O.ppo_rnd_backward_chunked(P, cfg, s_b, te_b, ti_b, y_b, adv_b, obs_b, old_b, mask, chunk=128)
        opt.step()
    for _ in range(warmup):
        step()
    t_remat[0] = 0.0
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return {"value": steps * n / dt, "unit": "samples/s", "cores": threads, "kind": "port",
            "sample": f"{steps} optimiser steps x {n} samples (the benched minibatch) of the cfg3 model: oracle port of agents.py:284-508 "
                      f"(torch fp32, dropout 0, {threads} threads), minibatch re-materialised from a {N}-sample rollout per step as the "
                      f"reference does (agents.py:288-301; cfg3 holds 16384), {warmup} warm-up steps; 'port' because /root/reference "
                      "does not exist on the GPU box -- tests/test_oracle_golden.py pins this port to the unmodified reference",
            "ms_per_step": dt / steps * 1e3, "ms_per_step_rematerialisation": t_remat[0] / steps * 1e3, "batch": n,
            "same_config": True}


def numerics_record(pk):
    """SURVEY 8a rows 12-15 at the cfg3 (128 envs) and cfg5 (1024 envs) rollout sizes: achieved HBM GB/s of the
    RunningMeanStd update and the observation normalisation for uint8 frames (what the device-resident rollout holds) and
    float32 (what the reference's numpy path holds), against the measured copy peak; GAE is launch-latency bound
    (<= 4.5 MB per call) and is reported as microseconds.  CUDA events around single launches, a 256 MB buffer rewritten
    between launches (L2 flush); bytes = algorithmic read + write of the call."""
    from eavit_b200 import ops
    F, Tn = 84 * 84, 128
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def timed(fn, reps=8):
        fn()
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / reps
    rows = []
    for E in (128, 1024):
        N = E * Tn
        mean = torch.zeros(F, dtype=torch.float64, device="cuda")
        var = torch.ones(F, dtype=torch.float64, device="cuda")
        cnt = torch.full((1,), 1e-4, dtype=torch.float64, device="cuda")
        for dt, sz, nm in ((torch.uint8, 1, "u8"), (torch.float32, 4, "f32")):
            x = (torch.rand(N, F, device="cuda") * 255).to(dt)
            out = torch.empty(N, F, dtype=torch.float32, device="cuda")
            for name, fn, nbytes in ((f"rms_update {nm}", lambda: ops.rms_update(x, mean, var, cnt), N * F * sz),
                                     (f"obs_normalize {nm}->f32", lambda: ops.obs_normalize(x, mean, var, out=out), N * F * (sz + 4))):
                ms = timed(fn)
                rows.append({"kernel": name, "envs": E, "us": ms * 1e3, "mbytes": nbytes / 1e6, "gbs": nbytes / ms / 1e6,
                             "frac_of_hbm_peak": nbytes / ms / 1e6 / pk["hbm_gbs"]})
            del x, out
        r = torch.rand(E, Tn, device="cuda")
        v = torch.randn(E, Tn + 1, device="cuda")
        done = (torch.rand(E, Tn, device="cuda") < 0.05).to(torch.uint8)
        rows.append({"kernel": "gae_f64 (numpy-promotion exact)", "envs": E, "us": timed(lambda: ops.gae_f64(r.double(), done, v, 0.999, 0.95, 0)) * 1e3,
                     "bound": "launch latency"})
        rows.append({"kernel": "gae_f32 warp-shuffle scan", "envs": E, "us": timed(lambda: ops.gae_f32(r, None, v, 0.99, 0.95)) * 1e3,
                     "bound": "launch latency"})
    del flush
    torch.cuda.empty_cache()
    return {"peak_gbs": pk["hbm_gbs"], "method": "CUDA events per call, 256 MB L2 flush between calls, bytes = algorithmic read + write", "rows": rows}


def side_agent(conf_name, E, overrides=None):
    """A second agent on another BASELINE config.  config.default_config is a module global that constructors read once, so
    the headline agent is unaffected by the reload."""
    import eavit_b200  # noqa: F401
    from eavit_b200 import agents, config, utils
    if conf_name is None:
        config.load_config(None, **(overrides or {}))
    else:
        config.load_config(os.path.join(ROOT, "configs", conf_name), **(overrides or {}))
    c = config.default_config
    N = E * T
    utils.set_seed(42)
    return agents.RNDAgent(84, A, utils.Env_action_space_type.DISCRETE, E, T, float(c["Gamma"]), GAE_Lambda=float(c["GAELambda"]),
                           learning_rate=float(c["LearningRate"]), ent_coef=float(c["Entropy"]), epoch=1, batch_size=N // int(c["MiniBatch"]),
                           ppo_eps=float(c["PPOEps"]), use_cuda=True, representation_lr_method="None",
                           device=f"cuda:{torch.cuda.current_device()}", logger=utils.Logger())


def side_record(agent, E, world, steps, flops_per_sample, label, barrier):
    """ms/step of `agent` (device-resident synthetic rollout of E envs x 128 steps, all-reduce included when world > 1)."""
    dev = agent.runtime().device
    N = E * T
    B = agent.batch_size
    R = dict(states=torch.randint(0, 256, (N, 4, 84, 84), dtype=torch.uint8, device=dev), te=torch.randn(N, device=dev),
             ti=torch.randn(N, device=dev), adv=torch.randn(N, device=dev), y=torch.randint(0, A, (N,), device=dev),
             obs=torch.randn(N, 1, 84, 84, device=dev).clamp_(-5, 5), old=torch.randn(N, A, device=dev))
    perm = torch.randperm(N, device=dev)
    mask = (torch.rand(B, device=dev) < 0.25).float()
    n_mb = N // B
    agent.runtime().sync()
    for i in range(3):
        agent.train_step(R, perm[B * (i % n_mb): B * (i % n_mb + 1)], mask)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        agent.train_step(R, perm[B * (i % n_mb): B * (i % n_mb + 1)], mask)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        t = torch.tensor([ms], device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t.item())
    rec = {"workload": label, "n_gpus": world, "envs_per_gpu": E, "minibatch_per_gpu": B, "steps": steps, "ms_per_step": ms,
           "value": B * world / (ms * 1e-3), "unit": "samples/s", "params": int(agent.runtime().store.numel)}
    if flops_per_sample:
        rec["model_tflops"] = flops_per_sample * B * world / (ms * 1e-3) / 1e12
        rec["model_tflops_per_gpu"] = rec["model_tflops"] / world
    del R
    return rec


def device_rollout_e2e(agent, E, barrier):
    """SURVEY 8f rows 1-2 end to end on the device: a filled DeviceRollout (uint8 frames written env-major at append time)
    -> finish() (reward filter, both GAE streams, advantage combine, obs statistics + normalisation) -> train_model with
    the CUDA tensors it returns (no host round trip) -> D2H of the update's loss terms."""
    from eavit_b200 import rollout, utils
    dev = agent.runtime().device
    ro = rollout.DeviceRollout(E, T, A, device=dev)
    g = torch.Generator(device=dev).manual_seed(3)
    ro.states.copy_(torch.randint(0, 256, ro.states.shape, dtype=torch.uint8, device=dev, generator=g))
    ro.next_obs.copy_(torch.randint(0, 256, ro.next_obs.shape, dtype=torch.uint8, device=dev, generator=g))
    ro.reward.copy_(torch.randn(E, T, device=dev, generator=g, dtype=torch.float64))
    ro.done.copy_((torch.rand(E, T, device=dev, generator=g) < 0.05).to(torch.uint8))
    ro.action.copy_(torch.randint(0, A, (E, T), device=dev, generator=g))
    ro.value_ext.copy_(torch.randn(E, T + 1, device=dev, generator=g)); ro.value_int.copy_(torch.randn(E, T + 1, device=dev, generator=g))
    ro.policy.copy_(torch.randn(E, T, A, device=dev, generator=g)); ro.int_reward.copy_(torch.rand(E, T, device=dev, generator=g))
    obs_rms, rew_rms = utils.RunningMeanStd(shape=(1, 1, 84, 84), usage="obs_rms"), utils.RunningMeanStd(usage="reward_rms")
    flt = utils.RewardForwardFilter(0.99)
    obs_rms.update(ro.next_obs.view(E * T, 1, 84, 84)[: 4 * E])
    ep = agent.epoch
    agent.epoch = 1
    agent.train_model(*ro.finish(obs_rms, rew_rms, flt, 0.999, 0.99, 0.95, 2.0, 1.0), 1)     # untimed warm-up update
    agent.epoch = ep
    barrier()
    t0 = time.perf_counter()
    args = ro.finish(obs_rms, rew_rms, flt, 0.999, 0.99, 0.95, 2.0, 1.0)
    torch.cuda.synchronize()
    t_fin = time.perf_counter() - t0
    agent.train_model(*args, 1)
    stats = agent.stats_summary()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return {"value": E * T * agent.epoch / dt, "unit": "samples/s", "seconds": dt, "finish_seconds": t_fin, "h2d_bytes_per_step": 0,
            "d2h_bytes_per_step": 16 * 4, "loss": stats.get("loss"),
            "call": "DeviceRollout.finish(...) -> RNDAgent.train_model(*cuda tensors): uint8 frames resident since append time"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-side", action="store_true", help="skip the cfg2 / cfg4 / numerics records")
    ap.add_argument("--profile-out", default=None, help="write the per-kernel CUDA-event table here (json)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return
        k = max(2, min(args.steps, 8))
        r = cpu_reference(k, max(1, min(args.warmup, 2)))
        line = {"impl": "reference", "metric": "PPO+RND update samples/sec", "value": r["value"], "unit": "samples/s",
                "n_gpus": args.gpus, "steps": k, "warmup": max(1, min(args.warmup, 2)), "ms_per_step": r["ms_per_step"],
                "steps_note": "bounded: at most 8 timed 512-sample optimiser steps (3-5 s each on the host cores)",
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD, "global_batch": 512},
                "cpu_baseline": {k2: r[k2] for k2 in ("value", "unit", "cores", "kind", "sample", "same_config", "ms_per_step_rematerialisation")},
                "e2e": {"value": r["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    assert torch.cuda.is_available(), "bench.py (impl b200) needs a CUDA device"
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from eavit_b200 import _lib, ops

    E = E_PER_GPU
    N = E * T
    B = N // MINI_BATCH
    agent = make_agent(E)
    upd_args, u8 = synth_update_args(E, seed=100 + rank)
    rt = agent.runtime()
    rt.sync()
    dev = rt.device
    R = agent.upload_rollout(*upd_args)
    perm = torch.from_numpy(np.random.default_rng(5).permutation(N)).to(dev)
    masks = (torch.rand(args.steps + args.warmup + 8, B) < 0.25).float().to(dev)
    n_mb = N // B

    def step(i):
        j = i % n_mb
        agent.train_step(R, perm[B * j: B * (j + 1)], masks[i % masks.shape[0]])

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(args.warmup + i)
    e1.record()
    barrier()
    launches = _lib.launch_count() - l0
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t.item())
    sampler.stop_flag = True
    value = args.steps * B * world / (ms * 1e-3)

    # ---- per-kernel attribution (CUDA events around every launch, separate pass) -> roofline of the dominant kernel
    pk, pk_kind = peaks()
    ops.profile_start()
    nprof = 3
    for i in range(nprof):
        step(i)
    table = ops.profile_stop()
    tot_ms = sum(v[1] for v in table.values())
    groups = {}
    for label, (n, tms, flops) in table.items():
        g = label.split(" ")[0]
        a = groups.setdefault(g, [0, 0.0, 0.0])
        a[0] += n; a[1] += tms; a[2] += flops * n
    # dominant kernel = largest share of the step among the dense-contraction kernels (every one carries its algorithmic FLOPs)
    dom_label, (dn, dms, dfl) = max(((l, v) for l, v in table.items() if v[2] > 0), key=lambda kv: kv[1][1])
    gemm = groups.get("gemm_bf16_tcgen05", [0, 0.0, 0.0])
    achieved = dfl / (dms / dn * 1e-3) / 1e12
    # DRAM bytes per launch of that kernel from the committed `ncu --set full` capture (profiles/r1_ncu_full_top_kernels.md)
    traffic = next((v for k, v in NCU_DRAM_BYTES_PER_LAUNCH.items() if dom_label.startswith(k)), None)
    roofline = {"bound": "tensor", "kernel": dom_label, "achieved": achieved, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                "frac": achieved / pk["bf16_tflops"], "traffic": traffic, "peak_source": f"{pk_kind} burst bf16 (kernel timed alone per launch)",
                "avg_launch_us": dms / dn * 1e3, "share_of_step": dms / tot_ms,
                "all_gemm": {"launches_per_step": gemm[0] / nprof, "share_of_step": gemm[1] / tot_ms,
                             "achieved_tflops": gemm[2] / (gemm[1] * 1e-3) / 1e12 if gemm[1] > 0 else None},
                "step_model_tflops": FLOPS_PER_SAMPLE * B / (ms / args.steps * 1e-3) / 1e12,
                "step_executed_tflops": FLOPS_PER_SAMPLE * executed_fraction() * B / (ms / args.steps * 1e-3) / 1e12,
                "executed_fraction_of_nominal_flops": executed_fraction(),
                "step_frac_of_sustained_peak": FLOPS_PER_SAMPLE * B / (ms / args.steps * 1e-3) / 1e12 / pk["bf16_tflops_sustained"],
                # every kernel family of the step (CUDA events around each launch, eager pass): launches, time, share, and the
                # achieved TFLOP/s of the dense contractions
                "per_kernel": [{"kernel": gname, "launches_per_step": a[0] / nprof, "ms_per_step": a[1] / nprof, "share": a[1] / tot_ms,
                                "tflops": (a[2] / (a[1] * 1e-3) / 1e12 if a[2] > 0 else None)}
                               for gname, a in sorted(groups.items(), key=lambda kv: -kv[1][1])[:14]]}
    # Both ceilings of the classical roofline for that kernel.  `bound` / `frac` above stay the tensor-pipe view (the kernel is a
    # dense contraction; same definition as round 1); at its arithmetic intensity the LOWER ceiling is the HBM one, so the
    # fraction of the binding ceiling is reported beside it.  Algorithmic bytes of the attention kernels at [T, 3*inner] bf16:
    # backward reads qkv, O, dO and writes dqkv (8 * T * inner * 2 B); forward reads qkv, writes O (4 * T * inner * 2 B).
    T_tok, inner = B * (196 + 197), 256
    alg_bytes = {"attention_bwd": 8 * T_tok * inner * 2, "attention_fwd": 4 * T_tok * inner * 2}
    ab = next((v for k, v in alg_bytes.items() if dom_label.startswith(k)), None)
    if ab is not None:
        t_launch = dms / dn * 1e-3
        ridge = pk["bf16_tflops"] * 1e12 / (pk["hbm_gbs"] * 1e9)
        ai = dfl / ab
        roofline["ceilings"] = {
            "algorithmic_bytes": ab, "arithmetic_intensity_flop_per_byte": ai, "ridge_flop_per_byte": ridge,
            "binding": "hbm" if ai < ridge else "tensor",
            "hbm": {"achieved": ab / t_launch / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ab / t_launch / 1e9 / pk["hbm_gbs"]},
            "tensor": {"achieved": achieved, "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": achieved / pk["bf16_tflops"]},
            "frac_of_binding_ceiling": max(ab / t_launch / 1e9 / pk["hbm_gbs"], achieved / pk["bf16_tflops"])}
    if args.profile_out and rank == 0:
        rows = sorted(((l, n / nprof, tms / nprof, fl) for l, (n, tms, fl) in table.items()), key=lambda r: -r[2])
        json.dump({"per_step": [{"kernel": l, "launches": n, "ms": tms, "tflops": (fl * n / (tms * 1e-3) / 1e12 if fl else None)}
                                for l, n, tms, fl in rows], "sum_ms": tot_ms / nprof}, open(args.profile_out, "w"), indent=1)

    # ---- secondary numbers: (M2) bare ViT fwd+bwd obs/s, and the same step with the shipped dropout keys (0.1)
    def timed_loop(fn, n):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(n):
            fn(i)
        b.record()
        barrier()
        t = a.elapsed_time(b)
        if world > 1:
            tt = torch.tensor([t], device=dev)
            torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
            t = float(tt.item())
        return t / n
    dpol = torch.randn(B, A, device=dev) * 1e-3
    dv = torch.randn(2 * B, device=dev) * 1e-3

    def vit_only(i):
        j = i % n_mb
        rt.ac_forward(R["states"], B, perm[B * j: B * (j + 1)])
        rt.ac_backward(dpol, dv)
    vit_only(0)
    vit_ms = timed_loop(vit_only, max(4, args.steps // 2))
    vit = {"metric": "ViT fwd+bwd obs/sec (both attention passes + heads)", "value": B * world / (vit_ms * 1e-3), "ms": vit_ms,
           "model_tflops": 6.36e9 * B / (vit_ms * 1e-3) / 1e12,
           "frac_of_sustained_bf16_peak": 6.36e9 * B / (vit_ms * 1e-3) / 1e12 / pk["bf16_tflops_sustained"]}
    c = rt.cfg
    saved = (c.dropout, c.emb_dropout, c.attn_dropout, c.act_dropout)
    c.dropout = c.emb_dropout = c.attn_dropout = c.act_dropout = 0.1          # the shipped expGlados3 keys
    for i in range(2):
        step(i)
    drop_ms = timed_loop(step, max(4, args.steps // 2))
    if args.profile_out and rank == 0:                 # per-kernel table of the dropout-0.1 step, next to the dropout-0 one
        ops.profile_start()
        step(0)
        dtab = ops.profile_stop()
        drows = sorted(((l, n, tms) for l, (n, tms, fl) in dtab.items()), key=lambda r: -r[2])
        json.dump({"per_step": [{"kernel": l, "launches": n, "ms": tms, "tflops": None} for l, n, tms in drows],
                   "sum_ms": sum(r[2] for r in drows)}, open(args.profile_out.replace(".json", "_dropout.json"), "w"), indent=1)
    c.dropout, c.emb_dropout, c.attn_dropout, c.act_dropout = saved
    with_dropout = {"dropout": 0.1, "value": B * world / (drop_ms * 1e-3), "unit": "samples/s", "ms_per_step": drop_ms}

    # ---- the same step over a long back-to-back loop (>= 200 steps ~ 1.6 s: clocks settle under the power cap)
    sus_n = max(200, args.steps)
    sus_ms = timed_loop(step, sus_n)
    sustained = {"steps": sus_n, "ms_per_step": sus_ms, "value": B * world / (sus_ms * 1e-3), "unit": "samples/s",
                 "step_executed_tflops": FLOPS_PER_SAMPLE * executed_fraction() * B / (sus_ms * 1e-3) / 1e12,
                 "step_executed_frac_of_sustained_peak": FLOPS_PER_SAMPLE * executed_fraction() * B / (sus_ms * 1e-3) / 1e12 / pk["bf16_tflops_sustained"]}

    # ---- end to end through the reference-facing call (host buffers in, stats out)
    def e2e_run(call_args, desc):
        agent.epoch = 1
        agent.train_model(*call_args, 1)               # untimed warm-up update: buffers of this argument signature, step-graph capture
        agent.epoch = EPOCH
        barrier()
        t0 = time.perf_counter()
        agent.train_model(*call_args, 1)
        stats = agent.stats_summary()                  # D2H read of the update's loss terms
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=dev)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            dt = float(t.item())
        n_steps = EPOCH * n_mb
        h2d = sum(a.nbytes for a in call_args)
        big = torch.from_numpy(call_args[0])
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        big.to(dev, non_blocking=True)
        torch.cuda.synchronize()
        return {"value": N * EPOCH * world / dt, "unit": "samples/s", "h2d_bytes_per_step": h2d / n_steps, "d2h_bytes_per_step": 16 * 4,
                "call": desc, "seconds": dt, "loss": stats.get("loss"), "host_buffers_pinned": bool(big.is_pinned()),
                "h2d_gb_per_s": call_args[0].nbytes / (time.perf_counter() - t1) / 1e9, "h2d_bytes_per_update": h2d}
    e2e = e2e_ref = e2e_dev = None
    if not args.no_e2e:
        # what the trainer holds: raw uint8 frames (train.py:585 appends them; the / 255 of train.py:854 happens in the patch
        # kernel, bit-identical) and the normalised next-obs as float32 (agents.py:298 converts to float32 anyway)
        def pin(a):
            t = torch.from_numpy(a)
            try:
                t = t.pin_memory()
            except Exception:
                pass
            return t.numpy()
        st, te_, ti_, y_, adv_, obs_, old_ = upd_args
        slim = (pin(u8), te_, ti_, y_, adv_, pin(obs_.astype(np.float32)), old_)
        e2e = e2e_run(slim, "RNDAgent.train_model(states uint8 frames, target_ext f64, target_int f64, y i64, adv f64, next_obs f32, "
                            "old_policy f32) -- numpy host buffers (pinned), one full update = 128 optimiser steps")
        e2e_ref = e2e_run(upd_args, "same call with the reference's own argument dtypes: states f32 (/255 on the host), next_obs f64")
        e2e_dev = device_rollout_e2e(agent, E, barrier)

    # ---- the other BASELINE configs, measured by the same command
    hg = cnn = None
    if not args.no_side:
        R.clear()                                        # free the headline rollout (3 GB) before the side agents allocate
        torch.cuda.empty_cache()
        E_hg = max(32, 256 // max(world, 2))            # cfg4: 256 envs over 2 / 4 GPUs; a single GPU runs the 2-GPU per-GPU shape
        a_hg = side_agent("vit_hg_explorative.conf", E_hg)
        hg = side_record(a_hg, E_hg, world, 8, 76.43e9, "cfg4: vit_hg HF-style ViT 1024/12L/16h/3072 (50-token sequences) RND agent, "
                         f"{E_hg} envs x 128 steps per GPU, all-reduce of the 150 M-parameter gradient included", barrier)
        del a_hg
        torch.cuda.empty_cache()
        if world == 1:
            a_cnn = side_agent(None, 64, {"ViT_implementation_type": 2, "extracted_feature_embedding_dim": 448})
            cnn = side_record(a_cnn, 64, 1, 50, None, "cfg2: original RND CNN actor-critic backbone (model.py:110-178), 64 envs x 128 steps, "
                              "minibatch 256", barrier)
            del a_cnn
            torch.cuda.empty_cache()

    in_sync = None
    if world > 1:
        # data-parallel invariant: same initial weights + all-reduced gradients -> bit-identical weights on every rank
        flat = rt.store.flat
        lo, hi = flat.clone(), flat.clone()
        torch.distributed.all_reduce(lo, op=torch.distributed.ReduceOp.MIN)
        torch.distributed.all_reduce(hi, op=torch.distributed.ReduceOp.MAX)
        in_sync = bool(torch.equal(lo, hi))
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    if rank != 0:
        return
    cpu = None if args.no_cpu else cpu_reference(3, 1)
    numerics = numerics_record(pk) if (world == 1 and not args.no_side) else None
    line = {"metric": "PPO+RND update samples/sec", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": B * world, "envs_per_gpu": E, "num_step": T, "parallelism": f"dp{world}",
                       "l2": "inputs larger than L2 (>= 5 GB of activations per step)", "timing": "CUDA events, max over ranks"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "gpu_launches_note": "kernels of libeavit_b200.so inside the timed region: direct launches + the kernel nodes of every replay of the captured step graph",
            "clocks": sampler.summary(),
            "vit_fwd_bwd": vit, "with_shipped_dropout": with_dropout, "weights_in_sync_across_ranks": in_sync,
            "sustained": sustained, "e2e_reference_dtypes": e2e_ref, "e2e_device_rollout": e2e_dev, "vit_hg_cfg4": hg,
            "cfg2_cnn_backbone": cnn, "numerics": numerics}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
