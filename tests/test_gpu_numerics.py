"""GPU parity: normalisation / GAE / reward-filter kernels (through the C ABI) vs the CPU oracle."""
import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from eavit_b200 import ops as _ops
    assert torch.cuda.is_available()
    return _ops


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("E,T", [(6, 16), (128, 128), (37, 200), (1024, 128), (1, 1)])
def test_gae_f64_bit_exact(ops, E, T):
    rng = np.random.default_rng(E * 1000 + T)
    rw = rng.normal(0, 1, (E, T)).clip(-1, 1)
    dn = rng.random((E, T)) < 0.05
    v = rng.normal(0, 1, (E, T + 1)).astype(np.float32)
    et, ea = O.make_train_data(rw, dn, v, 0.999, T, E)
    r, a = ops.gae_f64(dev(rw), dev(dn.astype(np.uint8)), dev(v), 0.999, 0.95, 0)
    assert np.array_equal(r.cpu().numpy(), et) and np.array_equal(a.cpu().numpy(), ea)
    ir = rng.random((E, T)).astype(np.float32)
    it, ia = O.make_train_data(ir, np.zeros_like(ir), v, 0.99, T, E)
    assert it.dtype == np.float64
    r, a = ops.gae_f64(dev(ir), None, dev(v), 0.99, 0.95, 1)
    assert np.array_equal(r.cpu().numpy(), it) and np.array_equal(a.cpu().numpy(), ia)


@pytest.mark.parametrize("E,T", [(6, 16), (128, 128), (33, 100), (1024, 128)])
def test_gae_f32_scan_within_1e5(ops, E, T):
    rng = np.random.default_rng(7)
    rw = rng.normal(0, 1, (E, T)).astype(np.float32)
    dn = rng.random((E, T)) < 0.05
    v = rng.normal(0, 1, (E, T + 1)).astype(np.float32)
    et, ea = O.make_train_data(rw.astype(np.float64), dn, v, 0.999, T, E)
    r, a = ops.gae_f32(dev(rw), dev(dn.astype(np.uint8)), dev(v), 0.999, 0.95)
    scale = np.abs(et).max()
    assert np.abs(r.cpu().numpy() - et).max() <= 1e-5 * scale       # north_star: 1e-5 for fp32 GAE
    assert np.abs(a.cpu().numpy() - ea).max() <= 1e-5 * scale


def test_axpby(ops):
    rng = np.random.default_rng(0)
    a, b = rng.normal(size=1000), rng.normal(size=1000)
    out = ops.axpby_f64(dev(a), dev(b), 1.0, 2.0)
    assert np.array_equal(out.cpu().numpy(), a * 1.0 + b * 2.0)       # train.py:767


@pytest.mark.parametrize("dtype", [np.uint8, np.float32, np.float64])
@pytest.mark.parametrize("N", [96, 2048, 5])
def test_rms_update_and_normalize(ops, dtype, N):
    rng = np.random.default_rng(N)
    F = 84 * 84
    x = rng.integers(0, 256, (N, 1, 84, 84), dtype=np.uint8)
    if dtype != np.uint8:
        x = (x.astype(np.float64) + rng.random(x.shape)).astype(dtype)
    rms = O.RunningMeanStd(shape=(1, 1, 84, 84))
    mean = torch.zeros(F, dtype=torch.float64, device="cuda")
    var = torch.ones(F, dtype=torch.float64, device="cuda")
    cnt = torch.full((1,), 1e-4, dtype=torch.float64, device="cuda")
    for rep in range(2):                               # second call exercises the Chan merge
        xx = x if rep == 0 else x[::-1].copy()
        rms.update(xx.astype(np.float64))
        ops.rms_update(dev(xx), mean, var, cnt)
    np.testing.assert_allclose(mean.cpu().numpy(), rms.mean.reshape(-1), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(var.cpu().numpy(), rms.var.reshape(-1), rtol=1e-10, atol=1e-10)
    assert abs(cnt.item() - rms.count) < 1e-9
    # normalise with identical stats -> float32 output identical to numpy's float64 math cast to f32
    m, v = dev(rms.mean.reshape(-1)), dev(rms.var.reshape(-1))
    ref = torch.FloatTensor(O.normalize_obs(x.astype(np.float64), rms)).numpy()
    got = ops.obs_normalize(dev(x), m, v).cpu().numpy()
    assert got.shape == ref.shape
    assert np.array_equal(got, ref)
    got16 = ops.obs_normalize(dev(x), m, v, out_dtype=torch.bfloat16).float().cpu().numpy()
    np.testing.assert_allclose(got16, ref, rtol=8e-3, atol=1e-6)


def test_rms_partial_merge_equals_update(ops):
    """Multi-GPU path: per-shard partial moments, summed, merged == single update over the whole batch."""
    rng = np.random.default_rng(1)
    F = 84 * 84
    x = rng.integers(0, 256, (64, F), dtype=np.uint8)
    st = lambda: (torch.zeros(F, dtype=torch.float64, device="cuda"), torch.ones(F, dtype=torch.float64, device="cuda"),
                  torch.full((1,), 1e-4, dtype=torch.float64, device="cuda"))
    m1, v1, c1 = st()
    ops.rms_update(dev(x), m1, v1, c1)
    m2, v2, c2 = st()
    s_a, q_a = ops.rms_partial(dev(x[:32]), m2)
    s_b, q_b = ops.rms_partial(dev(x[32:]), m2)
    ops.rms_merge(s_a + s_b, q_a + q_b, 64.0, m2, v2, c2)
    np.testing.assert_allclose(m2.cpu().numpy(), m1.cpu().numpy(), rtol=1e-13)
    np.testing.assert_allclose(v2.cpu().numpy(), v1.cpu().numpy(), rtol=1e-11)
    assert c1.item() == c2.item()


@pytest.mark.parametrize("E,T", [(6, 16), (128, 128), (1500, 32)])
def test_reward_filter_and_scale(ops, E, T):
    rng = np.random.default_rng(E)
    ir = rng.random((E, T)).astype(np.float32)
    filt, rrm = O.RewardForwardFilter(0.99), O.RunningMeanStd()
    rewems = torch.zeros(E, dtype=torch.float32, device="cuda")
    mean = torch.zeros(1, dtype=torch.float64, device="cuda")
    for rep in range(2):
        x = ir * (rep + 1)
        per_env = np.array([filt.update(r) for r in x.T])
        mom = ops.reward_filter(dev(x), rewems, rep > 0, 0.99).cpu().numpy()
        assert np.array_equal(rewems.cpu().numpy(), filt.rewems)            # float32 recurrence is bit-exact
        np.testing.assert_allclose(mom[0], np.mean(per_env.astype(np.float64)), rtol=1e-12)
        np.testing.assert_allclose(mom[1], np.var(per_env.astype(np.float64)), rtol=1e-9)
        np.testing.assert_allclose(mom[0], np.mean(per_env), rtol=1e-5)      # vs numpy's float32 pairwise mean
        assert mom[2] == T
        rrm.update_from_moments(np.mean(per_env), np.std(per_env) ** 2, len(per_env))
    var = torch.tensor([rrm.var], dtype=torch.float64, device="cuda")
    ref = ir.copy()
    ref /= np.sqrt(rrm.var)
    got = ops.scale_by_rsqrt_var(dev(ir), var).cpu().numpy()
    assert np.array_equal(got, ref)


def test_intrinsic_mse(ops):
    rng = np.random.default_rng(2)
    t, p = rng.normal(size=(77, 512)).astype(np.float32), rng.normal(size=(77, 512)).astype(np.float32)
    ref = (torch.tensor(t) - torch.tensor(p)).pow(2).mean(1).numpy()
    np.testing.assert_allclose(ops.intrinsic_mse(dev(t), dev(p)).cpu().numpy(), ref, rtol=1e-5)


@pytest.mark.parametrize("N,hw", [(8192, 84), (12345, 84), (16384, 84), (8200, 20), (300, 20), (4100, 12)])
def test_u8_rollout_size_paths(ops, N, hw):
    """uint8 frames at rollout size: the single-launch 16-pixels-per-lane RunningMeanStd update (ticketed last-CTA merge,
    utils.py:83-115) and the global-table normalisation (train.py:666 / :855) -- float64 statistics within round-off of
    numpy, normalised float32 output bit-identical; ragged row counts, F smaller than one column block."""
    rng = np.random.default_rng(N + hw)
    F = hw * hw
    x = rng.integers(0, 256, (N, 1, hw, hw), dtype=np.uint8)
    rms = O.RunningMeanStd(shape=(1, 1, hw, hw))
    mean = torch.zeros(F, dtype=torch.float64, device="cuda")
    var = torch.ones(F, dtype=torch.float64, device="cuda")
    cnt = torch.full((1,), 1e-4, dtype=torch.float64, device="cuda")
    xd = dev(x)
    for rep in range(3):                               # repeated calls: tickets must be back at zero, Chan merge chains
        xx = x if rep != 1 else (255 - x)
        rms.update(xx.astype(np.float64))
        ops.rms_update(xd if rep != 1 else 255 - xd, mean, var, cnt)
    np.testing.assert_allclose(mean.cpu().numpy(), rms.mean.reshape(-1), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(var.cpu().numpy(), rms.var.reshape(-1), rtol=1e-10, atol=1e-10)
    assert abs(cnt.item() - rms.count) < 1e-6
    # batch moments about a shift (the multi-GPU exchange) == numpy
    sh = dev(rms.mean.reshape(-1))
    s, q = ops.rms_partial(xd.view(N, F), sh)
    d = x.reshape(N, F).astype(np.float64) - rms.mean.reshape(-1)
    np.testing.assert_allclose(s.cpu().numpy(), d.sum(0), rtol=1e-9, atol=1e-6)
    np.testing.assert_allclose(q.cpu().numpy(), (d * d).sum(0), rtol=1e-11)
    m, v = dev(rms.mean.reshape(-1)), dev(rms.var.reshape(-1))
    ref = torch.FloatTensor(O.normalize_obs(x.astype(np.float64), rms)).numpy()
    got = ops.obs_normalize(xd, m, v).cpu().numpy()
    assert np.array_equal(got, ref)
    got16 = ops.obs_normalize(xd, m, v, out_dtype=torch.bfloat16)
    assert torch.equal(got16.cpu(), torch.from_numpy(ref).to(torch.bfloat16))
