"""CPU tests of the host side: config contract, drop-in parameter tree, ABI table, permutation / mask determinism."""
import os
from collections import OrderedDict
import re

import numpy as np
import pytest
import torch

from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_specs_match_header():
    """Every ctypes argument string in ops._SPECS matches the parameter list declared in include/eavit_b200.h."""
    from eavit_b200 import _lib, ops
    txt = re.sub(r"/\*.*?\*/", "", open(_lib.HEADER_PATH).read(), flags=re.S)
    for name, spec in ops._SPECS.items():
        m = re.search(r"\b" + name + r"\s*\((.*?)\)\s*;", txt, flags=re.S)
        assert m, name
        params = [p.strip() for p in m.group(1).split(",")]
        assert params[-1].endswith("stream")
        params = params[:-1]
        assert len(params) == len(spec), name
        for p, c in zip(params, spec):
            exp = "p" if "*" in p else ("u" if "unsigned long long" in p else "l" if "long long" in p else "i" if p.startswith("int")
                                        else "f" if p.startswith("float") else "d")
            assert exp == c, (name, p, c)
    declared = set(_lib.declared_symbols())
    assert set(ops._SPECS) <= declared


def test_config_contract():
    from eavit_b200 import config
    c = config.load_config(os.path.join(ROOT, "configs", "expGlados3_lucidrains_explorative.conf"))
    assert c["TrainMethod"] == "original_RND" and int(c["MiniBatch"]) == 32 and int(c["NumStep"]) == 128
    assert c.getboolean("ViTlucidrains_use_explorativeAttn") is True
    hp = config.HotPathConfig.from_config()
    assert (hp.impl, hp.dim, hp.depth, hp.heads, hp.dim_head, hp.mlp_dim, hp.patch, hp.n_patches, hp.patch_dim) == \
        ("lucidrains", 256, 3, 8, 32, 1024, 6, 196, 144)
    assert hp.dropout == 0.1          # shipped value; parity / bench runs override it to 0.0
    c = config.load_config(os.path.join(ROOT, "configs", "vit_hg_explorative.conf"))
    hp = config.HotPathConfig.from_config()
    assert (hp.impl, hp.dim, hp.depth, hp.heads, hp.dim_head, hp.n_patches, hp.ln_eps) == ("hg", 1024, 12, 16, 64, 49, 1e-12)
    config.load_config(None, ViTlucidrains_dropout=0.0)
    assert float(config.default_config["ViTlucidrains_dropout"]) == 0.0


@pytest.mark.parametrize("which", ["lucid", "cls", "hg"])
def test_state_dict_names_and_shapes_match_reference(which):
    """Parameter tree == the reference's state_dict (names + shapes pinned by tests/golden/make_golden.py, which
    load_state_dict(strict=True)'s the same dict into the reference classes)."""
    from eavit_b200 import config, model, utils
    cfg = {"lucid": O.OracleConfig(), "cls": O.OracleConfig(use_explorative=False),
           "hg": O.OracleConfig(impl="hg", patch=12, dim=128, depth=2, heads=2, dim_head=64, mlp_dim=256)}[which]
    if which == "hg":
        config.load_config(None, ViT_implementation_type=1, ViTHG_hidden_size=128, ViTHG_num_hidden_layers=2,
                           ViTHG_num_attention_heads=2, ViTHG_intermediate_size=256, extracted_feature_embedding_dim=128)
        impl = model.ViT_IMPLEMENTATION.HG_ViT
    else:
        config.load_config(None, ViTlucidrains_use_explorativeAttn=cfg.use_explorative)
        impl = model.ViT_IMPLEMENTATION.LUCIDRAINS_ViT
    ac = model.CnnActorCriticNetwork(84, 18, utils.Env_action_space_type.DISCRETE, False, ViT_implementation_type=impl)
    rnd = model.RNDModel(input_size=84, output_size=512, train_method="original_RND")
    got = {"model." + k: tuple(v.shape) for k, v in ac.state_dict().items()}
    got.update({"rnd." + k: tuple(v.shape) for k, v in rnd.state_dict().items()})
    assert got == O.param_shapes(cfg)
    assert not any(p.requires_grad for p in rnd.target.parameters())       # model.py:453-455
    assert all(p.requires_grad for p in rnd.predictor.parameters())
    config.load_config(None)


def test_cnn_backbone_constructor_follows_the_commented_reference_code():
    """BASELINE configs[1]: the CNN actor-critic that model.py:110-178 keeps as commented-out code.  The reference cannot
    construct it, so the check is against a line-by-line torch restatement of that block under the same seed: same
    parameter names / shapes as the oracle's table and bit-identical seeded initial weights."""
    import torch.nn as nn
    from torch.nn import init
    from eavit_b200 import config, model, utils
    config.load_config(None, ViT_implementation_type=2, extracted_feature_embedding_dim=448)
    utils.set_seed(42)
    ac = model.CnnActorCriticNetwork(84, 18, utils.Env_action_space_type.DISCRETE, False,
                                     ViT_implementation_type=model.ViT_IMPLEMENTATION.ORIGINAL_CNN)
    got = {"model." + k: tuple(v.shape) for k, v in ac.state_dict().items()}
    want = {k: v for k, v in O.param_shapes(O.OracleConfig(impl="cnn", dim=448)).items() if k.startswith("model.")}
    assert got == want
    utils.set_seed(42)
    D = 448
    feature = nn.Sequential(nn.Conv2d(4, 32, 8, 4), nn.ReLU(), nn.Conv2d(32, 64, 4, 2), nn.ReLU(), nn.Conv2d(64, 64, 3, 1), nn.ReLU(),
                            nn.Flatten(), nn.Linear(7 * 7 * 64, 256), nn.ReLU(), nn.Linear(256, D), nn.ReLU())        # :110-135
    actor = nn.Sequential(nn.Linear(D, D), nn.ReLU(), nn.Linear(D, 18))                                              # :137-141
    extra = nn.Sequential(nn.Linear(D, D), nn.ReLU())                                                                # :143-146
    c_ext, c_int = nn.Linear(D, 1), nn.Linear(D, 1)                                                                  # :148-149
    for m in (*feature, *actor, *extra, c_ext, c_int):              # `for p in self.modules()` order (:151-158)
        if isinstance(m, (nn.Conv2d, nn.Linear)):
            init.orthogonal_(m.weight, np.sqrt(2))
            m.bias.data.zero_()
    init.orthogonal_(c_ext.weight, 0.01); init.orthogonal_(c_int.weight, 0.01)                                       # :160-164
    for m in actor:
        if isinstance(m, nn.Linear):
            init.orthogonal_(m.weight, 0.01)                                                                         # :166-169
    for m in extra:
        if isinstance(m, nn.Linear):
            init.orthogonal_(m.weight, 0.1)                                                                          # :171-174
    ref = {**{"feature." + k: v for k, v in feature.state_dict().items()}, **{"actor." + k: v for k, v in actor.state_dict().items()},
           **{"extra_layer." + k: v for k, v in extra.state_dict().items()}, **{"critic_ext." + k: v for k, v in c_ext.state_dict().items()},
           **{"critic_int." + k: v for k, v in c_int.state_dict().items()}}
    sd = ac.state_dict()
    assert set(sd) == set(ref)
    for k in sd:
        assert torch.equal(sd[k], ref[k]), k
    config.load_config(None)


def test_seeded_init_matches_reference(golden_dir):
    """Constructors consume torch's RNG in the reference's order: same seed -> same initial weights (lucidrains)."""
    from eavit_b200 import config, model, utils
    G = np.load(os.path.join(golden_dir, "golden_init.npz"))
    config.load_config(None, ViTlucidrains_dropout=0.0, ViTlucidrains_emb_dropout=0.0)
    utils.set_seed(42)
    ac = model.CnnActorCriticNetwork(84, 18, utils.Env_action_space_type.DISCRETE, False)
    rnd = model.RNDModel(input_size=84, output_size=512, train_method="original_RND")
    sd = {"model." + k: v for k, v in ac.state_dict().items()}
    sd.update({"rnd." + k: v for k, v in rnd.state_dict().items()})
    for k, v in sd.items():
        a = v.double().reshape(-1).numpy()
        idx = np.linspace(0, a.size - 1, 64).astype(np.int64)
        d = np.concatenate([[np.sqrt((a * a).sum()), a.sum()], a[idx]])
        np.testing.assert_allclose(d, G[k], rtol=1e-6, atol=1e-7, err_msg=k)


def test_minibatch_permutation_and_mask_are_host_rng_bit_exact():
    """agents.py:270-285 / :336: the index stream is numpy MT19937 shuffles of ONE persistent arange, the RND mask is
    torch.rand on the CPU generator; the product draws them with the same calls (agents.train_model), so they are
    bit-identical by construction.  This pins the oracle's stream against a literal transcription."""
    N, B, epochs = 64, 16, 3
    np.random.seed(9)
    torch.manual_seed(9)
    sample_range = np.arange(N)
    ref_idx, ref_mask = [], []
    for _ in range(epochs):
        np.random.shuffle(sample_range)
        for j in range(N // B):
            ref_idx.append(sample_range[B * j:B * (j + 1)].copy())
            ref_mask.append((torch.rand(B) < 0.25).float())
    # the product's order of draws: all masks up front (torch stream), shuffles in the loop (numpy stream)
    np.random.seed(9)
    torch.manual_seed(9)
    masks = torch.stack([(torch.rand(B) < 0.25).float() for _ in range(epochs * (N // B))])
    sr = np.arange(N)
    got_idx = []
    for _ in range(epochs):
        np.random.shuffle(sr)
        for j in range(N // B):
            got_idx.append(sr[B * j:B * (j + 1)].copy())
    assert all(np.array_equal(a, b) for a, b in zip(ref_idx, got_idx))
    assert torch.equal(masks, torch.stack(ref_mask))
    assert sorted(np.concatenate(got_idx[: N // B]).tolist()) == list(range(N))      # each epoch covers every sample once


def test_no_cpu_fallback():
    """The product refuses to compute without CUDA instead of silently falling back."""
    from eavit_b200 import config, model, utils
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    config.load_config(None, ViTlucidrains_dropout=0.0, ViTlucidrains_emb_dropout=0.0)
    ac = model.CnnActorCriticNetwork(84, 18, utils.Env_action_space_type.DISCRETE, False)
    with pytest.raises(RuntimeError):
        ac(torch.zeros(1, 4, 84, 84))
    with pytest.raises(RuntimeError):
        utils.make_train_data(np.zeros((2, 4)), np.zeros((2, 4), dtype=bool), np.zeros((2, 5), dtype=np.float32), 0.99, 4, 2)


def test_torch_custom_ops_are_registered_and_cuda_only():
    """torch.ops.eavit_b200.* exist after import (torch.library) and refuse CPU tensors: no fallback path."""
    import eavit_b200  # noqa: F401
    for name in ("linear", "layer_norm", "attention", "gae", "intrinsic_mse", "obs_normalize"):
        assert hasattr(torch.ops.eavit_b200, name), name
    x = torch.randn(4, 8).bfloat16()
    with pytest.raises(NotImplementedError):
        torch.ops.eavit_b200.linear(x, x, None)


def test_rnd_gradient_slice_of_the_flat_store():
    """The early all-reduce covers exactly the predictor's contiguous block of the flat gradient, or nothing."""
    from eavit_b200.agents import RNDAgent

    class S:
        pass
    st = S(); st.offsets = {"model.a": 0, "model.b": 8, "rnd.predictor.0.weight": 16, "rnd.predictor.0.bias": 48}; st.numel = 56
    assert RNDAgent._rnd_grad_range(st) == (16, 56)
    st = S(); st.offsets = {"rnd.predictor.0.weight": 0, "rnd.predictor.0.bias": 32, "model.a": 40}; st.numel = 48
    assert RNDAgent._rnd_grad_range(st) == (0, 40)
    st = S(); st.offsets = {"rnd.predictor.0.weight": 0, "model.a": 32, "rnd.predictor.0.bias": 40}; st.numel = 48
    assert RNDAgent._rnd_grad_range(st) is None                 # interleaved: fall back to one all-reduce of everything
    st = S(); st.offsets = {"model.a": 0}; st.numel = 8
    assert RNDAgent._rnd_grad_range(st) is None


def test_gradient_exchange_ranges_partition_the_flat_buffer():
    """Overlapped gradient exchange: per-layer blocks + the predictor block + the remainder cover [0, numel) exactly once."""
    from eavit_b200.agents import RNDAgent
    from eavit_b200.engine import ParamStore
    assert RNDAgent._complement([], 40) == [(0, 40)]
    assert RNDAgent._complement([(8, 16), (24, 40)], 40) == [(0, 8), (16, 24)]
    assert RNDAgent._complement([(16, 24), (0, 8), (8, 16)], 24) == []
    shapes = O.param_shapes(O.OracleConfig())
    train = OrderedDict((k, v) for k, v in shapes.items() if not k.startswith("rnd.target."))
    st = ParamStore(train, "cpu", trainable=False)
    covered = []
    for li in range(3):
        names = [k for k in train if f"transformer.layers.{li}." in k]
        r = st.name_ranges(names)
        assert len(r) == 1, (li, r)                                  # a layer is ONE contiguous block: one all-reduce
        covered += r
    rnd = RNDAgent._rnd_grad_range(st)
    assert rnd is not None
    covered.append(rnd)
    rest = RNDAgent._complement(covered, st.numel)
    total = sorted(covered + rest)
    assert total[0][0] == 0 and total[-1][1] == st.numel
    assert all(a[1] == b[0] for a, b in zip(total, total[1:]))       # disjoint and gap-free
