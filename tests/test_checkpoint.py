"""Checkpoint format (train.py:883-960 writes, :198-238 reads): the pickled statistics objects of a
reference checkpoint restore this package's classes, a resumed run continues bit for bit, and a checkpoint written here
is a pickle of ``utils.RunningMeanStd`` / ``utils.RewardForwardFilter`` attribute bags that the reference restores.

Fixture: tests/golden/golden_ckpt_aux.{pt,npz}, written by the unmodified reference classes (make_golden_ckpt.py)."""
import io
import os
import pickletools
import sys
import types
import zipfile

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
PT = os.path.join(HERE, "golden", "golden_ckpt_aux.pt")
NPZ = os.path.join(HERE, "golden", "golden_ckpt_aux.npz")


def _pickle_globals(path):
    with zipfile.ZipFile(path) as z:
        name = [n for n in z.namelist() if n.endswith("data.pkl")][0]
        ops = list(pickletools.genops(z.read(name)))
    out, strings = set(), []
    for op, arg, _ in ops:
        if op.name in ("BINUNICODE", "SHORT_BINUNICODE", "UNICODE"):
            strings.append(arg)
        elif op.name == "STACK_GLOBAL":
            out.add((strings[-2], strings[-1]))
        elif op.name == "GLOBAL":
            out.add(tuple(arg.split(" ")))
    return out


class _StandInUtils:
    """A module named ``utils`` holding attribute-compatible stand-ins of the reference classes (utils.py:68-128): what
    ``torch.load`` needs on the reference side."""

    def __enter__(self):
        self.prev = sys.modules.get("utils")
        m = types.ModuleType("utils")

        class RunningMeanStd(object):
            pass

        class RewardForwardFilter(object):
            pass
        RunningMeanStd.__module__ = RewardForwardFilter.__module__ = "utils"
        RunningMeanStd.__qualname__, RewardForwardFilter.__qualname__ = "RunningMeanStd", "RewardForwardFilter"
        m.RunningMeanStd, m.RewardForwardFilter = RunningMeanStd, RewardForwardFilter
        sys.modules["utils"] = m
        return m

    def __exit__(self, *a):
        if self.prev is None:
            del sys.modules["utils"]
        else:
            sys.modules["utils"] = self.prev


def test_fixture_is_a_reference_pickle():
    g = _pickle_globals(PT)
    assert ("utils", "RunningMeanStd") in g and ("utils", "RewardForwardFilter") in g


def test_scalar_statistics_round_trip_in_reference_format(tmp_path):
    """CPU half: reward_rms (three host doubles) and the reward filter -> save_checkpoint -> plain torch.load on the
    'reference side' -> the reference's attribute names and values."""
    import eavit_b200  # noqa
    from eavit_b200 import utils
    rms = utils.RunningMeanStd(usage="reward_rms")
    rms.update_from_moments(0.25, 0.5, 16)
    filt = utils.RewardForwardFilter(0.99)
    filt.update(np.arange(4, dtype=np.float32))
    filt.update(np.ones(4, dtype=np.float32))
    p = str(tmp_path / "c.pt")
    utils.save_checkpoint({"reward_rms": rms, "discounted_reward": filt, "global_update": 7}, p)
    assert "utils" not in sys.modules or not hasattr(sys.modules["utils"], "__eavit_tmp__")
    g = _pickle_globals(p)
    assert ("utils", "RunningMeanStd") in g and ("utils", "RewardForwardFilter") in g
    assert not any(m.startswith("eavit_b200") or "explorative" in m for m, _ in g)      # nothing the reference cannot import
    with _StandInUtils() as ref_utils:
        ck = torch.load(p, weights_only=False)
        assert type(ck["reward_rms"]) is ref_utils.RunningMeanStd and type(ck["discounted_reward"]) is ref_utils.RewardForwardFilter
        assert set(ck["reward_rms"].__dict__) == {"usage", "mean", "var", "count", "train_method"}       # utils.py:72-78
        assert set(ck["discounted_reward"].__dict__) == {"rewems", "gamma"}                              # utils.py:119-121
        assert float(ck["reward_rms"].mean) == float(rms.mean) and float(ck["reward_rms"].var) == float(rms.var)
        assert ck["reward_rms"].count == rms.count and ck["reward_rms"].usage == "reward_rms"
        assert np.array_equal(ck["discounted_reward"].rewems, filt.rewems) and ck["discounted_reward"].gamma == 0.99
        assert ck["global_update"] == 7
    back = utils.load_checkpoint(p)                                  # and our own loader reads it back as our classes
    assert isinstance(back["reward_rms"], utils.RunningMeanStd) and isinstance(back["discounted_reward"], utils.RewardForwardFilter)
    assert back["reward_rms"].count == rms.count and np.array_equal(back["discounted_reward"].rewems, filt.rewems)


@pytest.mark.gpu
def test_reference_checkpoint_resumes(tmp_path):
    import eavit_b200  # noqa
    from eavit_b200 import utils
    gold = np.load(NPZ)
    ck = utils.load_checkpoint(PT)
    obs_rms, reward_rms, filt = ck["obs_rms"], ck["reward_rms"], ck["discounted_reward"]
    assert isinstance(obs_rms, utils.RunningMeanStd) and isinstance(reward_rms, utils.RunningMeanStd)
    assert isinstance(filt, utils.RewardForwardFilter) and ck["global_update"] == 3 and ck["global_step"] == 384
    assert obs_rms.mean.shape == (1, 1, 84, 84) and obs_rms.device_state()[0].is_cuda
    assert np.array_equal(obs_rms.mean, gold["mean"]) and np.array_equal(obs_rms.var, gold["var"]) and obs_rms.count == float(gold["count"])
    assert float(reward_rms.mean) == float(gold["r_mean"]) and float(reward_rms.var) == float(gold["r_var"])
    assert np.array_equal(filt.rewems, gold["rewems"])
    # one more update with the data the fixture script drew next (same PCG64 stream)
    rng = np.random.default_rng(77)
    rng.integers(0, 256, (96, 1, 84, 84)); rng.random((8, 16))
    x1 = rng.integers(0, 256, (64, 1, 84, 84)).astype(np.float64)
    r2 = rng.random((8, 16)).astype(np.float32)
    obs_rms.update(x1)
    per_step = np.array([filt.update(r2[:, t]) for t in range(16)])
    reward_rms.update_from_moments(np.mean(per_step), np.std(per_step) ** 2, len(per_step))
    # device statistics continue within float64 round-off (same bounds as test_gpu_numerics); the host paths are bit-exact
    np.testing.assert_allclose(obs_rms.mean, gold["mean1"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(obs_rms.var, gold["var1"], rtol=1e-10, atol=1e-10)
    assert abs(obs_rms.count - float(gold["count1"])) < 1e-9
    assert float(reward_rms.mean) == float(gold["r_mean1"]) and float(reward_rms.var) == float(gold["r_var1"])
    assert np.array_equal(filt.rewems, gold["rewems1"])
    # write it back in the reference's format and read it on the "reference side"
    p = str(tmp_path / "resume.pt")
    utils.save_checkpoint({**ck, "obs_rms": obs_rms, "reward_rms": reward_rms, "discounted_reward": filt}, p)
    with _StandInUtils() as ref_utils:
        ck2 = torch.load(p, weights_only=False)
        o2 = ck2["obs_rms"]
        assert type(o2) is ref_utils.RunningMeanStd and set(o2.__dict__) == {"usage", "mean", "var", "count", "train_method"}
        assert isinstance(o2.mean, np.ndarray) and o2.mean.dtype == np.float64 and o2.mean.shape == (1, 1, 84, 84)
        assert np.array_equal(o2.mean, obs_rms.mean) and np.array_equal(o2.var, obs_rms.var) and o2.count == obs_rms.count
        assert o2.usage == "obs_rms" and o2.train_method == "original_RND"
